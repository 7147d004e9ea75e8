"""The comparison helpers of oracle/parity.py (used by bench.py's parity block and by the large GPU differentials) checked on the
CPU: lookup phase B as oracle/parity.py computes it (vectorised + oracle/caller.py) against the C++ restatement in
tests/dropin/abi_over_oracle.cpp, which the end-to-end fixtures pin to the reference's own output files."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import parity  # noqa: E402
from oracle.bindings import Checker, MsaBatch, msa_to_numpy  # noqa: E402
from ploidyfrost_b200.synth import workload as wl  # noqa: E402


class SiteBatch(C.Structure):
    _fields_ = [("n_bubbles", C.c_uint32), ("reserved", C.c_uint32), ("site_off", C.POINTER(C.c_uint64)),
                ("status", C.POINTER(C.c_uint8)), ("n_class", C.POINTER(C.c_uint8)), ("cov_off", C.POINTER(C.c_uint64)),
                ("cov", C.POINTER(C.c_uint64))]


@pytest.fixture(scope="module")
def abi(tmp_path_factory):
    d = tmp_path_factory.mktemp("abi")
    so = os.path.join(str(d), "libabi_oracle.so")
    orc = os.path.join(ROOT, "oracle")
    if not os.path.exists(os.path.join(orc, "libpforacle.so")):
        subprocess.run(["make", "-C", orc, "oracle"], check=True, capture_output=True)
    subprocess.run(["g++", "-O1", "-std=c++14", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "dropin", "abi_over_oracle.cpp"), "-o", so, "-L", orc, "-lpforacle",
                    "-Wl,-rpath," + orc, "-lpthread"], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("n_hap,p_indel", [(4, 0.001), (3, 0.004)])
def test_expected_site_cov_matches_the_pinned_restatement(abi, tmp_path, n_hap, p_indel):
    k = 25
    w = wl.Workload(77 + n_hap, 120000, n_hap, p_snp=0.01, p_indel=p_indel)
    bb = w.bubbles(k, 0, 120000, 3000)
    haps = [w.haplotype(i) for i in range(n_hap)]
    prefix = str(tmp_path / "db")
    wl.write_db_numpy(prefix, haps, k, 12.6, 5, version=0x200, lut_prefix_len=9, sig_len=9, n_bins=64)
    w.close()
    ctx, db = C.c_void_p(), C.c_void_p()
    assert abi.pf_init(0, C.byref(ctx)) == 0
    assert abi.pf_kmc_open(ctx, prefix.encode(), C.byref(db)) == 0
    mb = MsaBatch()
    abi.pf_align.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    assert abi.pf_align(ctx, 2.0, -1.0, -3.0, bb.bases.ctypes.data, bb.seq_off.ctypes.data, bb.bubble_off.ctypes.data, bb.n_bubbles,
                        C.byref(mb)) == 0
    msa = msa_to_numpy(mb)
    sb = SiteBatch()
    skip = np.ascontiguousarray(bb.bubble_type.astype(np.uint8))
    abi.pf_site_cov.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    assert abi.pf_site_cov(db, 2, 1000, skip.ctypes.data, C.byref(sb)) == 0
    ns, nc = int(msa["var_off"][-1]), int(msa["cls_off"][-1])
    got = {"status": np.ctypeslib.as_array(sb.status, shape=(ns,)).copy(), "n_class": np.ctypeslib.as_array(sb.n_class, shape=(ns,)).copy(),
           "cov": np.ctypeslib.as_array(sb.cov, shape=(nc,)).copy()}
    orc = Checker("oracle")
    h = orc.kmc_open(prefix)
    lookup = lambda b, o: orc.kmc_counts(h, b, o, k, mode=1, use_read_api=False)
    checked, status, ncls, cov = parity.expected_site_cov(msa, skip, k, 2, 1000, lookup, max_general=10 ** 6)
    orc.kmc_close(h)
    assert checked.all()
    assert (msa["var_kind"] == 1).sum() > 20 and (status == parity.SITE_OK).sum() > 200
    assert parity.compare_site_cov(got, msa, checked, status, ncls, cov) == 0
    # and the comparison notices a difference
    got["cov"][np.flatnonzero(got["cov"])[0]] += 1
    assert parity.compare_site_cov(got, msa, checked, status, ncls, cov) == 1
    # compare_msa: identical results compare equal, a flipped class id is found
    m2 = {kk: (v.copy() if isinstance(v, np.ndarray) else v) for kk, v in msa.items()}
    assert not parity.compare_msa(msa, m2).any()
    m2["cls"][len(m2["cls"]) // 2] ^= 1
    assert parity.compare_msa(msa, m2).sum() == 1
    abi.pf_kmc_close(db)
