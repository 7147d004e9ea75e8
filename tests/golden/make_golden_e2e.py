#!/usr/bin/env python
"""Regenerates tests/golden/e2e/: the UNMODIFIED reference (oracle/_ref/PloidyFrost + oracle/_ref/Bifrost, built by
`make -C oracle ref_full` from /root/reference) run end to end, `-t 1`, on a small synthetic tetraploid.

  haplotypes (pfsynth, seeded)  ->  Bifrost build -r haps.fa -k 25 -i -d  ->  dbg.gfa
  canonical 25-mers of the haplotypes, count = 12 per haplotype copy + seeded jitter  ->  db.kmc_pre / db.kmc_suf (KMC1, p = 5)
  PloidyFrost -g dbg.gfa -d db -t 1 -l 2 -u 1000 -o P  ->  PloidyFrost_output/P_*.txt

Kept as fixtures: the KMC database, P_alignseq.txt (every bubble the reference aligned: VarId, strict/branching, entrance id, exit
id, aligned row), P_Unitig_Id.txt reduced to the unitigs that are an entrance or an exit, and the coverage / frequency files.
tests/test_cpu_golden.py and tests/test_gpu_e2e.py regenerate the reference's rows from these inputs with the oracle and with
the CUDA path.  Runs in the dev container only (needs /root/reference); deterministic.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ploidyfrost_b200.synth import kmcdb, workload as wl  # noqa: E402

K, LOW, UP = 25, 2, 1000
GENOME, HAPS, SEED = 36000, 4, 20261017


def main():
    ref = os.path.join(ROOT, "oracle", "_ref")
    for b in ("PloidyFrost", "Bifrost"):
        if not os.path.exists(os.path.join(ref, b)):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref_full"], check=True)
    out = os.path.join(HERE, "e2e")
    os.makedirs(out, exist_ok=True)
    with tempfile.TemporaryDirectory() as d:
        w = wl.Workload(SEED, GENOME, HAPS, p_snp=0.01, p_indel=0.002, n_threads=2)
        seqs = [bytes(w.haplotype(i)).decode() for i in range(HAPS)]
        w.close()
        with open(os.path.join(d, "haps.fa"), "w") as f:
            for i, s in enumerate(seqs):
                f.write(f">hap{i}\n{s}\n")
        u, c = kmcdb.count_canonical_kmers(seqs, K)
        rng = np.random.default_rng(SEED)
        cnt = (c * 12 + rng.integers(0, 7, len(c))).astype(np.uint64)
        kmcdb.write_kmc_db(os.path.join(d, "db"), u, cnt, K, version=0, lut_prefix_len=5, counter_size=2)
        subprocess.run([os.path.join(ref, "Bifrost"), "build", "-r", "haps.fa", "-k", str(K), "-i", "-d", "-o", "dbg", "-t", "1"],
                       cwd=d, check=True, capture_output=True)
        r = subprocess.run([os.path.join(ref, "PloidyFrost"), "-g", "dbg.gfa", "-d", "db", "-t", "1", "-l", str(LOW), "-u", str(UP),
                            "-o", "P"], cwd=d, check=True, capture_output=True, text=True)
        po = os.path.join(d, "PloidyFrost_output")
        # unitigs that are an entrance or an exit of an aligned bubble (their lengths enter VarDis)
        need = set()
        for ln in open(os.path.join(po, "P_alignseq.txt")):
            p = ln.split("\t")
            need.add(p[2]); need.add(p[3])
        with open(os.path.join(out, "P_Unitig_Id.txt"), "w") as f:
            for ln in open(os.path.join(po, "P_Unitig_Id.txt")):
                if ln.split("\t")[0] in need:
                    f.write(ln)
        for name in ["P_alignseq.txt"] + [f"P_{a}{b}.txt" for a in ("bi", "tri", "tetra", "penta") for b in ("cov", "fre")]:
            shutil.copy(os.path.join(po, name), os.path.join(out, name))
        for ext in (".kmc_pre", ".kmc_suf"):
            shutil.copy(os.path.join(d, "db" + ext), os.path.join(out, "db" + ext))
        meta = {"k": K, "low": LOW, "up": UP, "genome": GENOME, "haplotypes": HAPS, "seed": SEED, "db_kmers": int(len(u)),
                "reference_log_tail": r.stdout.strip().splitlines()[-2:]}
        json.dump(meta, open(os.path.join(out, "meta.json"), "w"), indent=1)
    print(json.dumps(meta, indent=1))
    subprocess.run("ls -la " + out, shell=True)


if __name__ == "__main__":
    main()
