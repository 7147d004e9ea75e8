"""CPU-side checks of the product: the C-ABI library loads and exports every declared symbol, fails loudly
without a GPU, and the device alignment state machines (compiled for the host by tests/hostemu) match the
oracle.  No CUDA compute is called here."""
import ctypes as C
import re
import os
import subprocess

import numpy as np
import pytest

from oracle.bindings import MsaBatch, flatten_bubbles, msa_to_numpy
from tests import gen
from tests.util import assert_msa_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ploidyfrost_b200 import build, capi
    build.build_library()
    lib = capi.load()
    hdr = open(os.path.join(ROOT, "include", "pf_gpu.h")).read()
    declared = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libpfgpu.so does not export {name}"
    assert declared == set(capi.EXPORTS)


def test_no_cpu_fallback_without_gpu():
    import torch
    from ploidyfrost_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PfError):
        capi.Context(0)


def _emu_align(emu, bubbles, M=2.0, D=-1.0, G=-3.0):
    emu.pfemu_align.restype = C.c_void_p
    emu.pfemu_align.argtypes = [C.c_double] * 3 + [C.c_void_p] * 3 + [C.c_uint32, C.c_int, C.POINTER(MsaBatch)]
    emu.pfemu_msa_free.argtypes = [C.c_void_p]
    emu.pfemu_set_limits.argtypes = [C.c_uint32] * 5
    bases, off, boff = flatten_bubbles(bubbles)
    mb = MsaBatch()
    h = emu.pfemu_align(M, D, G, bases.ctypes.data, off.ctypes.data, boff.ctypes.data, len(bubbles), 8, C.byref(mb))
    r = msa_to_numpy(mb)
    emu.pfemu_msa_free(h)
    return r


@pytest.mark.parametrize("seed,kw,sc", [
    (21, {}, {}), (22, dict(alphabet="AC"), {}), (23, dict(max_indel=4, max_snp=5), {}),
    (24, {}, dict(M=2.5, D=-1.5, G=-3.5)), (25, dict(alphabet="AC"), dict(M=3, D=-2, G=-1.5)),
    (26, dict(len_range=(100, 300), max_indel_len=40), {}), (27, dict(alphabet="A", len_range=(5, 30)), {}),
])
def test_device_state_machines_on_host_match_oracle(oracle, hostemu, seed, kw, sc):
    hostemu.pfemu_set_limits.argtypes = [C.c_uint32] * 5
    hostemu.pfemu_set_layout.argtypes = [C.c_uint32] * 3
    hostemu.pfemu_set_limits(16, 64, 64, 0, 0)
    bubbles = gen.random_bubbles(seed, 1500, **kw)
    a = oracle.align_bubbles(bubbles, n_threads=8, **sc)
    try:
        # (diagonal-major flags, contiguous area) = msa_warp_kernel; (row-major, lane-interleaved) = msa_lane_kernel
        # (2, 1, T) = the skewed layout of msa_cta_kernel with T emulated lanes
        for diag, lanes, arg in ((1, 1, seed), (0, 1, seed), (0, 32, seed), (2, 1, 7), (2, 1, 64)):
            hostemu.pfemu_set_layout(diag, lanes, arg)
            b = _emu_align(hostemu, bubbles, **sc)
            assert_msa_equal(a, b, bubbles, f"seed {seed} diag={diag} lanes={lanes} arg={arg}")
    finally:
        hostemu.pfemu_set_layout(0, 1, 0)


def test_small_capacity_reports_overflow_not_wrong_answers(oracle, hostemu):
    """With a tiny work area the machine must flag the bubble (status != 0), never return a different answer."""
    hostemu.pfemu_set_limits.argtypes = [C.c_uint32] * 5
    hostemu.pfemu_set_limits(16, 2, 2, 0, 0)
    try:
        bubbles = gen.random_bubbles(31, 1500, alphabet="AC")
        a = oracle.align_bubbles(bubbles, n_threads=8)
        b = _emu_align(hostemu, bubbles)
        from oracle.bindings import msa_bubble
        flagged = 0
        for i in range(len(bubbles)):
            y = msa_bubble(b, i)
            if y["status"] != 0:
                flagged += 1
                assert y["rows"] == []
            else:
                assert msa_bubble(a, i) == y
        assert flagged > 0
    finally:
        hostemu.pfemu_set_limits(16, 64, 64, 0, 0)


def test_bound_reference_binary_has_no_cpu_path(tmp_path):
    """oracle/_ref/PloidyFrost_gpu (integration/Makefile) on a machine without a GPU: the estimation phase stops with pf_init's
    error and a non-zero exit code -- it does not fall back to the reference's CPU code that is still linked into the binary."""
    import shutil
    import subprocess
    import torch
    from tests import e2e_rows
    gpu_bin = os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_gpu")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    if e2e_rows.reference_binaries() is None or not os.path.exists(gpu_bin):
        pytest.skip("oracle/_ref/PloidyFrost_gpu not built (make -C integration)")
    ref_dir = tmp_path / "ref"
    run_dir = tmp_path / "gpu"
    ref_dir.mkdir(); run_dir.mkdir()
    e2e_rows.run_reference_config0(str(ref_dir), genome=40000)
    for name in ("dbg.gfa", "db.kmc_pre", "db.kmc_suf"):
        shutil.copy(ref_dir / name, run_dir / name)
    for threads in ("1", "4"):
        r = subprocess.run([gpu_bin, "-g", "dbg.gfa", "-d", "db", "-t", threads, "-l", "2", "-u", "1000", "-o", "P"], cwd=run_dir,
                           capture_output=True, text=True)
        assert r.returncode != 0
        assert "no CPU fallback" in r.stdout
        assert not os.path.exists(run_dir / "PloidyFrost_output" / "P_bicov.txt")


def _build_caller_over_oracle(tmp_path):
    exe = os.path.join(str(tmp_path), "caller_over_oracle")
    orc = os.path.join(ROOT, "oracle")
    if not os.path.exists(os.path.join(orc, "libpforacle.so")):
        subprocess.run(["make", "-C", orc, "oracle"], check=True, capture_output=True)
    subprocess.run(["g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "dropin", "caller_test.cpp"),
                    os.path.join(ROOT, "tests", "dropin", "abi_over_oracle.cpp"), "-o", exe, "-L", orc, "-lpforacle", "-Wl,-rpath," + orc,
                    "-lpthread"], check=True)
    return exe


def test_batched_caller_host_logic_writes_the_reference_files(tmp_path):
    """The HOST side of include/pf_caller.hpp (gating, ordering, strict-bubble arithmetic, VarDis, row text) with the device calls
    played by the CPU oracle (tests/dropin/abi_over_oracle.cpp, test infrastructure): fed with the bubbles of tests/golden/e2e it
    writes the unmodified reference's `-t 1` files byte for byte; in the `-t N` dialect the same rows with 0-based ids and
    P_allele_frequency grouped per bubble (CDBG.cpp:2056, :2162, :2550)."""
    import filecmp
    from collections import Counter
    exe = _build_caller_over_oracle(tmp_path)
    fx = os.path.join(ROOT, "tests", "golden", "e2e")
    names = ["P_alignseq.txt"] + [f"P_{a}{w}.txt" for a in ("bi", "tri", "tetra", "penta") for w in ("cov", "fre")]   # the fixture's files
    one = tmp_path / "t1"
    one.mkdir()
    r = subprocess.run([exe, fx, str(one), "2", "1000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for n in names:
        assert filecmp.cmp(os.path.join(fx, n), one / n, shallow=False), n
    # P_allele_frequency.txt is not in the fixture (tests/test_gpu_integration.py compares it with a live reference run): -t 1
    # writes every site, so it holds at least the lines of the four frequency files
    fre_lines = Counter(ln for a in ("bi", "tri", "tetra", "penta") for ln in open(one / f"P_{a}fre.txt"))
    assert not (fre_lines - Counter(open(one / "P_allele_frequency.txt")))
    mt = tmp_path / "mt"
    mt.mkdir()
    r = subprocess.run([exe, fx, str(mt), "2", "1000", "mt"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr

    def shift_id(line, col):
        p = line.split("\t")
        p[col] = str(int(p[col]) - 1)
        return "\t".join(p)
    assert [shift_id(ln, 0) for ln in open(one / "P_alignseq.txt")] == list(open(mt / "P_alignseq.txt"))
    for a in ("bi", "tri", "tetra", "penta"):
        assert [shift_id(ln, -4) for ln in open(one / f"P_{a}cov.txt")] == list(open(mt / f"P_{a}cov.txt"))
        assert filecmp.cmp(one / f"P_{a}fre.txt", mt / f"P_{a}fre.txt", shallow=False)
    all_mt = list(open(mt / "P_allele_frequency.txt"))
    assert not (Counter(all_mt) - Counter(open(one / "P_allele_frequency.txt")))
    n_fre = sum(len(list(open(mt / f"P_{a}fre.txt"))) for a in ("bi", "tri", "tetra"))
    assert n_fre <= len(all_mt) <= n_fre + len(list(open(mt / "P_pentafre.txt")))


@pytest.mark.parametrize("kw", [dict(haplotypes=4, p_indel=0.003, depth=60),                               # 3-8 branches per bubble
                                dict(haplotypes=4, p_indel=0.003, depth=60, low=10, up=40),              # tight gate: dropped sites
                                dict(haplotypes=2, p_indel=0.01, p_snp=0.03, depth=30, low=2, up=25)])   # variant-dense, low ceiling
def test_batched_caller_host_logic_on_a_live_reference_run(tmp_path, kw):
    """Same stand-in, bubbles of a fresh run of the unmodified reference: all ten files of the `-t 1` run, P_allele_frequency.txt
    included, byte for byte -- also under gates (-l / -u) that make the reference skip sites."""
    import filecmp
    import shutil
    from tests import e2e_rows
    if e2e_rows.reference_binaries() is None:
        pytest.skip("oracle/_ref/PloidyFrost not built (make -C oracle ref_full)")
    out, dbp = e2e_rows.run_reference_config0(str(tmp_path), genome=200000, **kw)
    for ext in (".kmc_pre", ".kmc_suf"):
        shutil.copy(dbp + ext, os.path.join(out, "db" + ext))
    exe = _build_caller_over_oracle(tmp_path)
    got = tmp_path / "got"
    got.mkdir()
    r = subprocess.run([exe, out, str(got), str(kw.get("low", 2)), str(kw.get("up", 1000))], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    names = sorted(os.listdir(got))
    assert len(names) == 10
    for n in names:
        assert filecmp.cmp(os.path.join(out, n), got / n, shallow=False), n
    # the rows formatted by 5 host threads (contiguous ranges joined in order): the same bytes
    got5 = tmp_path / "got5"
    got5.mkdir()
    r = subprocess.run([exe, out, str(got5), str(kw.get("low", 2)), str(kw.get("up", 1000)), "t1", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for n in names:
        assert filecmp.cmp(got / n, got5 / n, shallow=False), n
    assert os.path.getsize(got / "P_bicov.txt") > 10000


def _hostcheck_binary():
    """oracle/_ref/PloidyFrost_hostcheck: integration/ploidy_estimation_gpu.cpp inside the unmodified reference, linked against the
    oracle stand-in instead of libpfgpu.so (integration/Makefile `hostcheck`; dev container only, it needs the reference's objects)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_hostcheck")
    if not os.path.exists(exe) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "integration"), "hostcheck"], check=False, capture_output=True)
    return exe if os.path.exists(exe) else None


@pytest.mark.parametrize("kw", [dict(haplotypes=2, p_indel=0.001, depth=30),                              # BASELINE configs[0], scaled
                                dict(haplotypes=4, p_indel=0.003, depth=60, low=10, up=40),              # a third of the bubbles gated out
                                dict(haplotypes=2, p_indel=0.01, p_snp=0.03, depth=30, low=2, up=25)])
def test_bound_reference_binary_walk_and_host_logic(tmp_path, kw):
    """The reference-side binding end to end on the CPU: same binary as PloidyFrost_gpu except that the device calls are played by
    the oracle.  Its graph walk (which bubbles, which order, ids, sizes, branch strings, sort keys), the caller's host logic and the
    writers must reproduce every file of the unmodified binary byte for byte, and its closing statistics."""
    import filecmp
    import shutil
    from tests import e2e_rows
    hc = _hostcheck_binary()
    if e2e_rows.reference_binaries() is None or hc is None:
        pytest.skip("oracle/_ref/PloidyFrost_hostcheck not built (make -C integration hostcheck, dev container)")
    ref_dir, run_dir = tmp_path / "ref", tmp_path / "hc"
    ref_dir.mkdir(); run_dir.mkdir()
    low, up = kw.get("low", 2), kw.get("up", 1000)
    out, _ = e2e_rows.run_reference_config0(str(ref_dir), genome=200000, **kw)
    for name in ("dbg.gfa", "db.kmc_pre", "db.kmc_suf"):
        shutil.copy(ref_dir / name, run_dir / name)
    cmd = ["-g", "dbg.gfa", "-d", "db", "-t", "1", "-l", str(low), "-u", str(up), "-o", "P"]
    r = subprocess.run([hc] + cmd, cwd=run_dir, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    names = sorted(n for n in os.listdir(out) if n.startswith("P_"))
    assert len(names) == 12
    for n in names:
        assert filecmp.cmp(os.path.join(out, n), run_dir / "PloidyFrost_output" / n, shallow=False), n
    ref_stdout = subprocess.run([e2e_rows.reference_binaries()[0]] + cmd, cwd=ref_dir, capture_output=True, text=True).stdout

    def closing(text):
        return [ln for ln in text.split("\n") if "Alleles in SuperBubbles" in ln or "Average Coverage" in ln]
    assert closing(r.stdout) == closing(ref_stdout) and len(closing(r.stdout)) == 2
    # bubbles beyond a device limit go through the host aligner (the reference's own SeqAlign in this binary) and the host-built site
    # k-mers of include/pf_caller.hpp: force every other aligned bubble down that path -- same bytes
    r2 = subprocess.run([hc] + cmd, cwd=run_dir, capture_output=True, text=True, env=dict(os.environ, PF_CALLER_FORCE_HOST="2"))
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    m = [ln for ln in r2.stdout.split("\n") if "by the host aligner" in ln]
    assert m and " 0 by the host aligner" not in m[0]
    for n in names:
        assert filecmp.cmp(os.path.join(out, n), run_dir / "PloidyFrost_output" / n, shallow=False), n + " (host aligner path)"
    # -t 4: the reference's worker threads against the binding's single pass, as multisets (ids and order are schedule-dependent)
    cmd[5] = "4"
    subprocess.run([e2e_rows.reference_binaries()[0]] + cmd, cwd=ref_dir, check=True, capture_output=True)
    r = subprocess.run([hc] + cmd, cwd=run_dir, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want, _ = e2e_rows.thread_dialect_view(out)
    got, ids = e2e_rows.thread_dialect_view(str(run_dir / "PloidyFrost_output"))
    assert ids == list(range(len(ids)))
    for key in want:
        assert got[key] == want[key], key


def test_host_site_kmers_match_the_pinned_restatement(tmp_path):
    """include/pf_caller.hpp: host_site_kmers (the site k-mers of the few bubbles that go through the host aligner) against
    oracle/caller.py::site_kmers -- the restatement of CDBG.cpp:2331-2473 that is pinned on the reference's own end-to-end files --
    on random gapped alignments: SNP / indel sites, before / after an earlier indel site, k-mers that reach past either end."""
    import random
    from oracle import caller
    src = r'''
#include "pf_caller.hpp"
#include <iostream>
int main() {
    size_t n_cases; std::cin >> n_cases;
    for (size_t t = 0; t < n_cases; t++) {
        size_t n, c, k, ni; int indel;
        std::cin >> n >> c >> k >> indel >> ni;
        std::vector<std::string> rows(n), out;
        for (auto &r : rows) std::cin >> r;
        if (!pfdropin::host_site_kmers(rows, c, k, indel != 0, ni, out)) { std::cout << "UNDEFINED\n"; continue; }
        for (size_t r = 0; r < n; r++) std::cout << out[r] << (r + 1 < n ? " " : "\n");
    }
}
'''
    (tmp_path / "t.cpp").write_text(src)
    exe = str(tmp_path / "t")
    subprocess.run(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), str(tmp_path / "t.cpp"), "-o", exe,
                    "-Wl,--unresolved-symbols=ignore-all", "-pthread"], check=True)
    rng = random.Random(7)
    cases, want = [], []
    while len(cases) < 3000:
        n, L, k = rng.randint(2, 5), rng.randint(30, 70), rng.choice([9, 15, 25])
        base = [rng.choice("ACGT") for _ in range(L)]
        rows = []
        for _ in range(n):
            r = list(base)
            for _ in range(rng.randint(0, 3)):
                r[rng.randrange(L)] = rng.choice("ACGT")
            for _ in range(rng.randint(0, 2)):                   # a gap run
                a = rng.randrange(L)
                for x in range(a, min(L, a + rng.randint(1, 4))):
                    r[x] = "-"
            rows.append("".join(r))
        c, is_indel, ni = rng.randrange(L), rng.random() < 0.5, rng.choice([0, 0, 1, 2])
        try:
            exp = caller.site_kmers(rows, c, k, is_indel, ni)
            if any(len(x) != k for x in exp):
                exp = None                                       # a slice that ran off the row: the reference would throw / misbehave
        except (AssertionError, IndexError):
            exp = None
        cases.append(f"{n} {c} {k} {int(is_indel)} {ni} " + " ".join(rows))
        want.append(exp)
    out = subprocess.run([exe], input=f"{len(cases)}\n" + "\n".join(cases) + "\n", capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(cases)
    n_defined = 0
    for line, exp, case in zip(out, want, cases):
        if exp is None:
            continue          # outside the reference's own domain: the header may report UNDEFINED or any string, no row is written from it
        n_defined += 1
        assert line.split(" ") == exp, case
    assert n_defined > 1000
