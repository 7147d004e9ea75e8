"""CPU-side checks of the product: the C-ABI library loads and exports every declared symbol, fails loudly
without a GPU, and the device alignment state machines (compiled for the host by tests/hostemu) match the
oracle.  No CUDA compute is called here."""
import ctypes as C
import re
import os

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
        for diag, lanes in ((1, 1), (0, 1), (0, 32)):
            hostemu.pfemu_set_layout(diag, lanes, seed)
            b = _emu_align(hostemu, bubbles, **sc)
            assert_msa_equal(a, b, bubbles, f"seed {seed} diag={diag} lanes={lanes}")
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
