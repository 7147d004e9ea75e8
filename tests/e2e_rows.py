"""TEST INFRASTRUCTURE.  Regenerates the rows PloidyFrost writes for the bubbles it aligned (coverage / frequency files, `-t 1`
dialect) from: the bubbles' raw branch strings, a SequenceAlignment implementation, a KMC lookup implementation.  The logic is a
restatement of the reference's per-bubble caller -- strict bubbles CDBG.cpp:1186-1330 (= :1998-2189), branching bubbles
:1440-1660 (= :2190-2575); the per-site pieces (site k-mers, class coverages, VarDis) live in oracle/caller.py -- used to pin the oracle AND the CUDA path against the unmodified
reference's own output files (tests/golden/e2e, made by tests/golden/make_golden_e2e.py)."""
from __future__ import annotations

import json
import os

import numpy as np

from oracle.bindings import flatten_bubbles, flatten_seqs, msa_bubble
from tests.refrun import colored_thread_dialect_view, reference_binaries, thread_dialect_view, unitig_seq  # noqa: F401
from oracle.caller import class_coverage, fmt, site_kmers, site_outcome, var_distance  # noqa: F401  (the CPU restatement)

HERE = os.path.dirname(os.path.abspath(__file__))
E2E = os.path.join(HERE, "golden", "e2e")


def parse_alignseq(path):
    """-> [dict(var_id, strict, ent, exit, rows)] in file order (consecutive lines of one VarId are one bubble)."""
    out = []
    for ln in open(path):
        p = ln.rstrip("\n").split("\t")
        if out and out[-1]["var_id"] == int(p[0]):
            out[-1]["rows"].append(p[4])
        else:
            out.append(dict(var_id=int(p[0]), strict=p[1] == "1", ent=int(p[2]), exit=int(p[3]), rows=[p[4]]))
    return out


def parse_cov_files(d):
    """-> {var_id: [row string, ...]} over the bi/tri/tetra/penta coverage files."""
    rows = {}
    for a in ("bi", "tri", "tetra", "penta"):
        for ln in open(os.path.join(d, f"P_{a}cov.txt")):
            p = ln.rstrip("\n").split("\t")
            rows.setdefault(int(p[-4]), []).append(ln.rstrip("\n"))   # ... type, indelLen, VarId, VarNum, VarDis, ''
    return rows


def branching_site_plan(bubbles, aligned, k):
    """Every variable column of every branching bubble with its site k-mers: [(bubble index, site index, column, is_indel,
    indel sites seen so far incl. this one, offset into the k-mer list)], [k-mer strings]."""
    plan, kmers = [], []
    for bi, (b, m) in enumerate(zip(bubbles, aligned)):
        if not m["rows"] or b["strict"]:
            continue
        n_ind = 0
        for i, c in enumerate(sorted(m["partition"].keys())):
            is_ind = c in m["indel_pos"]
            km = site_kmers(m["rows"], c, k, is_ind, n_ind)
            if is_ind:
                n_ind += 1
            plan.append((bi, i, c, is_ind, n_ind, len(kmers)))
            kmers.extend(km)
    return plan, kmers


def regenerate(bubbles, unitig_len, k, low, up, align_fn, cov_fn, kmer_count_fn, site_cov_fn=None):
    """bubbles: parse_alignseq() records.  align_fn(flat bubbles) -> MSA batch (numpy dict); cov_fn(bases, off) -> pf_cov records
    (readCov, 'as written else reverse complement'); kmer_count_fn(list of k-mer strings) -> (counts, found).
    site_cov_fn(msa, bubble index, site index) -> class coverages or None replaces the host-side site k-mer path when given
    (the device implementation of lookup phase B).  Returns ({var_id: [cov rows]}, {var_id: [frequencies]}, aligned)."""
    raw = [[r.replace("-", "") for r in b["rows"]] for b in bubbles]
    msa = align_fn(*flatten_bubbles(raw))
    aligned = [msa_bubble(msa, i) for i in range(len(bubbles))]
    bases, off = flatten_seqs([s for b in raw for s in b])
    cov = cov_fn(bases, off)                               # lookup-A: every branch string
    plan, kmers = branching_site_plan(bubbles, aligned, k)
    if site_cov_fn is None:
        counts, found = kmer_count_fn(kmers) if kmers else (np.zeros(0, np.uint32), np.zeros(0, np.uint8))
    cov_rows, fre_rows = {}, {}
    s0 = 0
    pi = 0
    for bi, (b, m) in enumerate(zip(bubbles, aligned)):
        nrow = len(b["rows"])
        vid = b["var_id"]
        if m["rows"]:
            var_site = sorted(m["partition"].keys())
            ent, ext = unitig_len[b["ent"]], unitig_len[b["exit"]]
            if b["strict"]:
                means = [float(cov["sum"][s0 + j]) / float(cov["n_kmers"][s0 + j]) for j in range(nrow)]
                total = 0.0
                for x in means:
                    total += x
                n_ind = 0
                for i, c in enumerate(var_site):
                    part = m["partition"][c]
                    tc = [0.0] * max(part)
                    for j, cl in enumerate(part):
                        tc[cl - 1] += means[j]
                    il = 0
                    if c in m["indel_pos"]:
                        n_ind += 1
                        il = m["indel_len"][n_ind - 1]
                    row = "".join(fmt(x) + "\t" for x in tc) + f"1\t{il}\t{vid}\t{len(var_site)}\t{var_distance(i, var_site, ent, ext)}\t"
                    cov_rows.setdefault(vid, []).append(row)
                    fre_rows.setdefault(vid, []).extend(fmt(x / total) for x in tc)
            else:
                while pi < len(plan) and plan[pi][0] == bi:
                    _, i, c, is_ind, n_ind, k0 = plan[pi]
                    pi += 1
                    if site_cov_fn is not None:
                        tc = site_cov_fn(msa, bi, i)
                    else:
                        tc = class_coverage(m["partition"][c], kmers, k0, counts, found, low, up)
                    if tc is None:
                        continue
                    total = 0.0
                    for x in tc:
                        total += x
                    il = m["indel_len"][n_ind - 1] if is_ind else 0
                    row = "".join(fmt(x) + "\t" for x in tc) + f"0\t{il}\t{vid}\t{len(var_site)}\t{var_distance(i, var_site, ent, ext)}\t"
                    cov_rows.setdefault(vid, []).append(row)
                    fre_rows.setdefault(vid, []).extend(fmt(x / total) for x in tc)
        s0 += nrow
    return cov_rows, fre_rows, aligned


def load_fixture(d=None):
    d = d or E2E
    meta = json.load(open(os.path.join(d, "meta.json")))
    bubbles = parse_alignseq(os.path.join(d, "P_alignseq.txt"))
    useq = {}
    for ln in open(os.path.join(d, "P_Unitig_Id.txt")):
        i, s = ln.rstrip("\n").split("\t")
        useq[int(i)] = s
    return meta, bubbles, useq, parse_cov_files(d)


def check_against_reference(align_fn, cov_fn, kmer_count_fn, site_cov_fn=None, fixture_dir=None):
    """Asserts that (a) SequenceAlignment of the raw branch strings reproduces the reference's aligned rows for every bubble and
    (b) the regenerated coverage rows equal the reference's files row for row.  Returns (#rows, #branching bubbles)."""
    meta, bubbles, useq, ref_rows = load_fixture(fixture_dir)
    k = meta["k"]
    ulen = {i: len(s) for i, s in useq.items()}             # UnitigMap::size is the unitig length in bases
    cov_rows, fre_rows, aligned = regenerate(bubbles, ulen, k, meta["low"], meta["up"], align_fn, cov_fn, kmer_count_fn, site_cov_fn)
    for b, m in zip(bubbles, aligned):
        assert m["rows"] == b["rows"], f"VarId {b['var_id']}: aligned rows differ from the reference's"
    n = 0
    for vid, rows in ref_rows.items():
        mine = sorted(cov_rows.get(vid, []))
        assert mine == sorted(rows), f"VarId {vid}:\n reference {sorted(rows)}\n ours      {mine}"
        n += len(rows)
    assert set(cov_rows) == set(ref_rows)
    return n, sum(1 for b in bubbles if not b["strict"])


def device_site_cov_hook(db, meta, bubbles):
    """site_cov_fn for check_against_reference() that takes lookup phase B from the device (pf_site_cov over the rows pf_align left
    in HBM).  Returns (hook, state); state["sc"] holds the raw pf_site_cov result after the first call."""
    from ploidyfrost_b200 import capi
    state = {}

    def site_cov(msa, bi, i):
        if "sc" not in state:
            skip = np.array([1 if b["strict"] else 0 for b in bubbles], np.uint8)
            state["sc"] = db.site_cov(meta["low"], meta["up"], skip)
            state["nrows"] = msa["n_rows"]
        sc = state["sc"]
        v = int(sc["site_off"][bi]) + i
        st = int(sc["status"][v])
        assert st in (capi.SITE_OK, capi.SITE_DROPPED), f"bubble {bi} site {i}: status {st}"
        if st == capi.SITE_DROPPED:
            return None
        c0 = int(sc["cov_off"][bi]) + i * int(state["nrows"][bi])
        return [float(x) for x in sc["cov"][c0:c0 + int(sc["n_class"][v])]]
    return site_cov, state


# ---- BASELINE configs[0]: the reference run here and now ----------------------------------------------------------------------


def run_reference_config0(workdir, genome=200000, depth=30, read_len=150, k=25, low=2, up=1000, seed=20261017, threads=1, haplotypes=2,
                          p_snp=0.01, p_indel=0.001):
    """BASELINE configs[0], scaled by `genome`: synthetic diploid (1 % SNP, 0.1 % indel), `depth`x error-free reads of random
    strand, Bifrost graph of the reads, KMC database = canonical k-mer counts of the reads; the unmodified PloidyFrost is run on
    it and its output directory is returned in the shape load_fixture() reads."""
    import subprocess
    from ploidyfrost_b200.synth import kmcdb, workload as wl
    pf, bf = reference_binaries()
    w = wl.Workload(seed, genome, haplotypes, p_snp=p_snp, p_indel=p_indel, n_threads=2)
    haps = [bytes(w.haplotype(i)).decode() for i in range(haplotypes)]
    w.close()
    rng = np.random.default_rng(seed)
    comp = str.maketrans("ACGT", "TGCA")
    reads = []
    for h in haps:
        n = depth * len(h) // (haplotypes * read_len)
        for a in rng.integers(0, len(h) - read_len, n):
            r = h[a:a + read_len]
            reads.append(r.translate(comp)[::-1] if rng.random() < 0.5 else r)
    with open(os.path.join(workdir, "reads.fa"), "w") as f:
        for i, r in enumerate(reads):
            f.write(f">r{i}\n{r}\n")
    u, c = kmcdb.count_canonical_kmers(reads, k)
    kmcdb.write_kmc_db(os.path.join(workdir, "db"), u, c.astype(np.uint64), k, version=0x200, lut_prefix_len=5, counter_size=2, n_bins=64,
                       sig_len=9)
    subprocess.run([bf, "build", "-r", "reads.fa", "-k", str(k), "-i", "-d", "-o", "dbg", "-t", "4"], cwd=workdir, check=True,
                   capture_output=True)
    subprocess.run([pf, "-g", "dbg.gfa", "-d", "db", "-t", str(threads), "-l", str(low), "-u", str(up), "-o", "P"], cwd=workdir, check=True,
                   capture_output=True)
    out = os.path.join(workdir, "PloidyFrost_output")
    json.dump({"k": k, "low": low, "up": up, "genome": genome, "db_kmers": int(len(u))}, open(os.path.join(out, "meta.json"), "w"))
    return out, os.path.join(workdir, "db")


# ---- the `-t N` files, schedule-independent ------------------------------------------------------------------------------------




# ---- coloured graphs (BASELINE configs[3]; CCDBG.cpp:538 / :2759) --------------------------------------------------------------
def make_colored_inputs(workdir, genome=120000, n_samples=3, depth=20, read_len=150, k=25, seed=20261017, haplotypes=4, p_snp=0.01,
                        p_indel=0.002, low=2, up=1000):
    """n_samples read sets over haplotype subsets of one synthetic polyploid (sample s holds the haplotypes h with (h + s) % 3 != 0,
    always at least two), `Bifrost build -c` over them (one colour per read file), one KMC database per sample, the database list
    file `dbs.txt` and the thresholds file `cov.txt` (`low\\tup` per line, Main.cpp:416-447).  Returns the per-sample database prefixes."""
    import subprocess
    from ploidyfrost_b200.synth import kmcdb, workload as wl
    pf, bf = reference_binaries()
    w = wl.Workload(seed, genome, haplotypes, p_snp=p_snp, p_indel=p_indel, n_threads=2)
    haps = [bytes(w.haplotype(i)).decode() for i in range(haplotypes)]
    w.close()
    rng = np.random.default_rng(seed)
    comp = str.maketrans("ACGT", "TGCA")
    prefixes, files = [], []
    for s in range(n_samples):
        mine = [h for h in range(haplotypes) if (h + s) % 3 != 0]
        if len(mine) < 2:
            mine = list(range(haplotypes))[:2]
        reads = []
        for h in mine:
            hs = haps[h]
            n = depth * len(hs) // read_len
            for a in rng.integers(0, len(hs) - read_len, n):
                r = hs[a:a + read_len]
                reads.append(r.translate(comp)[::-1] if rng.random() < 0.5 else r)
        fn = f"reads{s}.fa"
        with open(os.path.join(workdir, fn), "w") as f:
            for i, r in enumerate(reads):
                f.write(f">s{s}r{i}\n{r}\n")
        files.append(fn)
        u, c = kmcdb.count_canonical_kmers(reads, k)
        kmcdb.write_kmc_db(os.path.join(workdir, f"db{s}"), u, c.astype(np.uint64), k, version=0x200 if s % 2 == 0 else 0, lut_prefix_len=5,
                           counter_size=2, n_bins=64, sig_len=9)
        prefixes.append(f"db{s}")
    open(os.path.join(workdir, "dbs.txt"), "w").write("".join(p + "\n" for p in prefixes))
    open(os.path.join(workdir, "cov.txt"), "w").write("".join(f"{low}\t{up}\n" for _ in prefixes))
    cmd = [bf, "build", "-c", "-k", str(k), "-i", "-d", "-o", "dbg", "-t", "4"]
    for fn in files:
        cmd += ["-r", fn]
    subprocess.run(cmd, cwd=workdir, check=True, capture_output=True)
    return prefixes


def run_reference_colored(workdir, threads=1, binary=None, **kw):
    """The unmodified PloidyFrost (or `binary`) on the coloured inputs of make_colored_inputs; returns its output directory."""
    import subprocess
    if not os.path.exists(os.path.join(workdir, "dbg.gfa")):
        make_colored_inputs(workdir, **kw)
    pf, _ = reference_binaries()
    r = subprocess.run([binary or pf, "-g", "dbg.gfa", "-f", "dbg.bfg_colors", "-d", "dbs.txt", "-C", "cov.txt", "-t", str(threads), "-o", "P"],
                       cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return os.path.join(workdir, "PloidyFrost_output")

