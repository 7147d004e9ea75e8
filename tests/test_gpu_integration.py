"""-m gpu: the drop-in, end to end.  `oracle/_ref/PloidyFrost_gpu` is the UNMODIFIED reference with one definition replaced
(integration/ploidy_estimation_gpu.cpp: CDBG::ploidyEstimation_ptr hands its bubbles to libpfgpu.so through include/pf_caller.hpp;
built by `make -C integration` in the dev container, travels with the snapshot).  It and the unmodified `PloidyFrost` are run
with the same command line on the same Bifrost graph and KMC database; every output file must be byte-identical."""
import filecmp
import os
import shutil
import subprocess

import pytest

from tests import e2e_rows

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BIN = os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_gpu")


@pytest.mark.parametrize("hap,genome,p_indel,p_snp,low,up", [(2, 300000, 0.001, 0.01, 2, 1000), (4, 300000, 0.003, 0.01, 2, 1000),
                                                             (4, 300000, 0.003, 0.01, 10, 40)])   # tight gate: dropped bubbles and sites
def test_patched_reference_binary_writes_identical_files(tmp_path, hap, genome, p_indel, p_snp, low, up):
    if e2e_rows.reference_binaries() is None or not os.path.exists(GPU_BIN):
        pytest.skip("oracle/_ref/PloidyFrost_gpu not built (make -C integration in the dev container)")
    ref_dir = tmp_path / "ref"
    gpu_dir = tmp_path / "gpu"
    ref_dir.mkdir(); gpu_dir.mkdir()
    out, dbp = e2e_rows.run_reference_config0(str(ref_dir), genome=genome, haplotypes=hap, p_indel=p_indel, p_snp=p_snp, depth=15 * hap,
                                              low=low, up=up)
    for name in ("dbg.gfa", "db.kmc_pre", "db.kmc_suf"):
        shutil.copy(ref_dir / name, gpu_dir / name)
    r = subprocess.run([GPU_BIN, "-g", "dbg.gfa", "-d", "db", "-t", "1", "-l", str(low), "-u", str(up), "-o", "P"], cwd=gpu_dir,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = gpu_dir / "PloidyFrost_output"
    names = sorted(n for n in os.listdir(out) if n.startswith("P_"))
    assert len(names) >= 12
    for n in names:
        assert filecmp.cmp(os.path.join(out, n), got / n, shallow=False), f"{n} differs from the unmodified reference's file"
    assert os.path.getsize(got / "P_bicov.txt") > 10000
    if hap == 4 and low == 2:
        # bubbles beyond a device limit are aligned by the reference's own SeqAlign inside this binary and their site k-mers are
        # built on the host (include/pf_caller.hpp, HostMsa): every third aligned bubble forced down that path -- same bytes
        r = subprocess.run([GPU_BIN, "-g", "dbg.gfa", "-d", "db", "-t", "1", "-l", str(low), "-u", str(up), "-o", "P"], cwd=gpu_dir,
                           capture_output=True, text=True, env=dict(os.environ, PF_CALLER_FORCE_HOST="3"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "by the host aligner" in r.stdout and " 0 by the host aligner" not in r.stdout
        for n in names:
            assert filecmp.cmp(os.path.join(out, n), got / n, shallow=False), f"{n} differs (host aligner path)"


def test_patched_reference_binary_thread_dialect(tmp_path):
    """`-t 4`: the reference's worker threads against ours (the device); files equal as multisets, ours with ids 0..n-1 in order."""
    if e2e_rows.reference_binaries() is None or not os.path.exists(GPU_BIN):
        pytest.skip("oracle/_ref/PloidyFrost_gpu not built (make -C integration in the dev container)")
    ref_dir = tmp_path / "ref"
    gpu_dir = tmp_path / "gpu"
    ref_dir.mkdir(); gpu_dir.mkdir()
    out, dbp = e2e_rows.run_reference_config0(str(ref_dir), genome=300000, haplotypes=4, p_indel=0.003, depth=60, threads=4)
    for name in ("dbg.gfa", "db.kmc_pre", "db.kmc_suf"):
        shutil.copy(ref_dir / name, gpu_dir / name)
    r = subprocess.run([GPU_BIN, "-g", "dbg.gfa", "-d", "db", "-t", "4", "-l", "2", "-u", "1000", "-o", "P"], cwd=gpu_dir,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want, _ = e2e_rows.thread_dialect_view(out)
    got, ids = e2e_rows.thread_dialect_view(str(gpu_dir / "PloidyFrost_output"))
    assert ids == list(range(len(ids))) and len(ids) > 1000
    for key in want:
        assert got[key] == want[key], f"{key}: differs from the unmodified reference's -t 4 run as a multiset"
