"""-m gpu: the coloured drop-in, end to end (BASELINE configs[3]).  `oracle/_ref/PloidyFrost_gpu` -- the UNMODIFIED reference with
CCDBG::ploidyEstimation_ptr / _multithread_ptr replaced by integration/ploidy_estimation_colored_gpu.cpp over libpfgpu.so -- and the
unmodified `PloidyFrost` run with the same command line on the same `Bifrost build -c` graph, per-sample KMC databases (KMC1 and
KMC2 layouts mixed) and `-C` thresholds; `-t 1`: every output file byte-identical; `-t 4`: equal as multisets."""
import filecmp
import os
import subprocess

import pytest

from tests import e2e_rows
from tests.test_cpu_colored import ARGS, link_inputs

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BIN = os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_gpu")


def need_binaries():
    if e2e_rows.reference_binaries() is None or not os.path.exists(GPU_BIN):
        pytest.skip("oracle/_ref/PloidyFrost_gpu not built (make -C integration in the dev container)")


@pytest.mark.parametrize("kw", [dict(genome=120000, n_samples=3, haplotypes=4, p_indel=0.003, low=12, up=45),
                                dict(genome=100000, n_samples=8, haplotypes=6, p_indel=0.004, low=3, up=800, depth=12)])
def test_colored_binary_t1_files_identical(tmp_path, kw):
    need_binaries()
    src = str(tmp_path / "ref")
    os.makedirs(src)
    want = e2e_rows.run_reference_colored(src, threads=1, **kw)
    gpu = str(tmp_path / "gpu")
    link_inputs(src, gpu)
    r = subprocess.run([GPU_BIN] + ARGS + ["-t", "1"], cwd=gpu, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = os.path.join(gpu, "PloidyFrost_output")
    names = sorted(n for n in os.listdir(want) if n.startswith("P_"))
    assert len(names) >= 12
    for n in names:
        assert filecmp.cmp(os.path.join(want, n), os.path.join(got, n), shallow=False), f"{n} differs from the unmodified reference's file"
    assert os.path.getsize(os.path.join(got, "P_bicov.txt")) > 5000
    if kw["n_samples"] == 3:   # the host aligner path (bubbles beyond a device limit), forced for every other aligned bubble
        r = subprocess.run([GPU_BIN] + ARGS + ["-t", "1"], cwd=gpu, capture_output=True, text=True, env=dict(os.environ, PF_CALLER_FORCE_HOST="2"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "by the host aligner" in r.stdout and " 0 by the host aligner" not in r.stdout
        for n in names:
            assert filecmp.cmp(os.path.join(want, n), os.path.join(got, n), shallow=False), f"{n} differs (host aligner path)"


def test_colored_binary_thread_dialect(tmp_path):
    need_binaries()
    src = str(tmp_path / "ref")
    os.makedirs(src)
    want_dir = e2e_rows.run_reference_colored(src, threads=4, genome=120000, n_samples=3, haplotypes=4, p_indel=0.003, low=2, up=1000)
    want, _ = e2e_rows.colored_thread_dialect_view(want_dir)
    gpu = str(tmp_path / "gpu")
    link_inputs(src, gpu)
    r = subprocess.run([GPU_BIN] + ARGS + ["-t", "4"], cwd=gpu, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got, ids = e2e_rows.colored_thread_dialect_view(os.path.join(gpu, "PloidyFrost_output"))
    assert ids == list(range(len(ids))) and len(ids) > 300
    for key in want:
        assert got[key] == want[key], f"{key}: differs from the unmodified reference's -t 4 run as a multiset"
