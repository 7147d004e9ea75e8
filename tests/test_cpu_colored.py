"""CPU pin of the COLOURED binding's host logic (BASELINE configs[3]; CCDBG.cpp:538 `-t N`, :2759 `-t 1`).

`oracle/_ref/PloidyFrost_hostcheck` is the reference binary with CCDBG::ploidyEstimation_ptr / _multithread_ptr replaced by
integration/ploidy_estimation_colored_gpu.cpp + include/pf_caller_colored.hpp, the device calls played by the CPU oracle
(tests/dropin/abi_over_oracle.cpp -- test infrastructure).  It and the unmodified `PloidyFrost` run on the same `Bifrost build -c`
graph, per-sample KMC databases and `-C` thresholds file; `-t 1` files must be byte-identical, `-t 4` files equal as multisets.
The same comparison with libpfgpu.so instead of the oracle is tests/test_gpu_colored.py."""
import filecmp
import os
import subprocess

import pytest

from tests import e2e_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTCHECK = os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_hostcheck")
ARGS = ["-g", "dbg.gfa", "-f", "dbg.bfg_colors", "-d", "dbs.txt", "-C", "cov.txt", "-o", "P"]


def link_inputs(src, dst):
    os.makedirs(dst)
    for f in os.listdir(src):
        if f.startswith("db") or f == "cov.txt":
            os.symlink(os.path.join(src, f), os.path.join(dst, f))


@pytest.fixture(scope="module")
def colored_inputs(tmp_path_factory):
    if e2e_rows.reference_binaries() is None or not os.path.exists(HOSTCHECK):
        pytest.skip("oracle/_ref/PloidyFrost_hostcheck not built (make -C integration hostcheck in the dev container)")
    d = str(tmp_path_factory.mktemp("colored"))
    e2e_rows.make_colored_inputs(d, genome=60000, n_samples=3, haplotypes=4, p_indel=0.003, low=12, up=45)   # a gate that drops bubbles and sites
    return d


def test_colored_binding_t1_files_identical(colored_inputs, tmp_path):
    want = e2e_rows.run_reference_colored(colored_inputs, threads=1)
    hc = str(tmp_path / "hc")
    link_inputs(colored_inputs, hc)
    r = subprocess.run([HOSTCHECK] + ARGS + ["-t", "1"], cwd=hc, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = os.path.join(hc, "PloidyFrost_output")
    names = sorted(n for n in os.listdir(want) if n.startswith("P_"))
    assert len(names) >= 12
    for n in names:
        assert filecmp.cmp(os.path.join(want, n), os.path.join(got, n), shallow=False), f"{n} differs from the unmodified reference's file"
    assert os.path.getsize(os.path.join(got, "P_bicov.txt")) > 3000
    # bubbles beyond a device limit go through the host aligner (the reference's own SeqAlign in this binary) and host-built site
    # k-mers: force every other aligned bubble down that path -- same bytes
    hc2 = str(tmp_path / "hc2")
    link_inputs(colored_inputs, hc2)
    r2 = subprocess.run([HOSTCHECK] + ARGS + ["-t", "1"], cwd=hc2, capture_output=True, text=True, env=dict(os.environ, PF_CALLER_FORCE_HOST="2"))
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert "by the host aligner" in r2.stdout and " 0 by the host aligner" not in r2.stdout
    for n in names:
        assert filecmp.cmp(os.path.join(want, n), os.path.join(hc2, "PloidyFrost_output", n), shallow=False), f"{n} differs (host aligner path)"
    # the average entrance coverage line of the console (CCDBG.cpp:1441) is part of the contract too
    ref = subprocess.run([e2e_rows.reference_binaries()[0]] + ARGS + ["-t", "1"], cwd=colored_inputs, capture_output=True, text=True)
    line = [ln for ln in ref.stdout.splitlines() if "Average Coverage" in ln]
    assert line and line[0] in r.stdout


def test_colored_binding_thread_dialect(colored_inputs, tmp_path):
    want_dir = e2e_rows.run_reference_colored(colored_inputs, threads=4)
    want, _ = e2e_rows.colored_thread_dialect_view(want_dir)
    hc = str(tmp_path / "hc")
    link_inputs(colored_inputs, hc)
    r = subprocess.run([HOSTCHECK] + ARGS + ["-t", "4"], cwd=hc, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got, ids = e2e_rows.colored_thread_dialect_view(os.path.join(hc, "PloidyFrost_output"))
    assert ids == list(range(len(ids))) and len(ids) > 100
    for key in want:
        assert got[key] == want[key], f"{key}: differs from the unmodified reference's -t 4 run as a multiset"


def test_cramer_v_matches_reference_formula():
    """computeCramerVCoefficient (CCDBG.cpp:330) on literal vectors, through a tiny C++ driver over the header."""
    import tempfile
    src = r'''
#include "pf_caller_colored.hpp"
#include <cstdio>
int main() {
    std::vector<double> a{10, 20, 0}, b{20, 10, 0}, c{0, 0, 5}, z{0, 0, 0};
    std::printf("%.17g %.17g %.17g %.17g\n", pfdropin::ColoredBubbleCaller::cramer_v(a, b), pfdropin::ColoredBubbleCaller::cramer_v(a, a),
                pfdropin::ColoredBubbleCaller::cramer_v(a, c), pfdropin::ColoredBubbleCaller::cramer_v(z, c));
}
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(src)
        subprocess.run(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.cpp"), "-o", os.path.join(d, "t"),
                        "-Wl,--unresolved-symbols=ignore-all", "-pthread"], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    v = [float(x) for x in out]
    # chi^2 of the 2 x 2 table [[10, 20], [20, 10]] = 4 * (5^2 / 15) = 6.666..., n = 60 -> sqrt(1/9)
    assert abs(v[0] - 1.0 / 3.0) < 1e-15 and v[1] == 0.0 and abs(v[2] - 1.0) < 1e-15 and v[3] == 0.0
