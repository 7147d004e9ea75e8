"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE -- may be imported only from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs, never from the product package.

  Checker("oracle") -> oracle/libpforacle.so   our own restatement (oracle/pf_oracle.cpp), symbols pforc_*
  Checker("ref")    -> oracle/_ref/libpfref.so the unmodified reference + oracle/ref_shim.cpp, symbols pfref_*

Both export the same call shapes, so one wrapper serves both.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)


class KmcInfo(C.Structure):
    _fields_ = [("kmer_length", C.c_uint32), ("mode", C.c_uint32), ("counter_size", C.c_uint32),
                ("lut_prefix_length", C.c_uint32), ("signature_len", C.c_uint32), ("min_count", C.c_uint32),
                ("max_count", C.c_uint64), ("total_kmers", C.c_uint64), ("both_strands", C.c_uint32),
                ("kmc_version", C.c_uint32), ("n_bins", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class Cov(C.Structure):
    _fields_ = [("sum", C.c_uint64), ("min", C.c_uint32), ("n_kmers", C.c_uint32),
                ("first_missing", C.c_int32), ("first_outside", C.c_int32)]


COV_DTYPE = np.dtype([("sum", "<u8"), ("min", "<u4"), ("n_kmers", "<u4"),
                      ("first_missing", "<i4"), ("first_outside", "<i4")])


class MsaBatch(C.Structure):
    _fields_ = [("n_bubbles", C.c_uint32), ("reserved", C.c_uint32), ("status", i32p), ("n_rows", u32p),
                ("aln_len", u32p), ("rows_off", u64p), ("rows", C.POINTER(C.c_char)), ("var_off", u64p),
                ("var_col", u32p), ("var_kind", u8p), ("cls_off", u64p), ("cls", u16p), ("ilen_off", u64p),
                ("ilen", u32p)]


def _np_from(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()


def msa_to_numpy(mb: MsaBatch) -> dict:
    """Deep-copies a pf_msa_batch_t view into a dict of numpy arrays (canonical comparison form)."""
    n = mb.n_bubbles
    out = {"n_bubbles": n}
    out["status"] = _np_from(mb.status, n, np.int32)
    out["n_rows"] = _np_from(mb.n_rows, n, np.uint32)
    out["aln_len"] = _np_from(mb.aln_len, n, np.uint32)
    for name, data, dt in (("rows", mb.rows, np.uint8), ("var", None, None), ("cls", mb.cls, np.uint16),
                           ("ilen", mb.ilen, np.uint32)):
        off = _np_from(getattr(mb, name + "_off"), n + 1, np.uint64)
        out[name + "_off"] = off
        total = int(off[-1]) if n else 0
        if name == "var":
            out["var_col"] = _np_from(mb.var_col, total, np.uint32)
            out["var_kind"] = _np_from(mb.var_kind, total, np.uint8)
        else:
            out[name] = _np_from(data, total, dt)
    return out


def msa_bubble(m: dict, b: int) -> dict:
    """One bubble of a msa_to_numpy() dict in the reference's own output shape."""
    nr, L = int(m["n_rows"][b]), int(m["aln_len"][b])
    r0 = int(m["rows_off"][b])
    rows = [bytes(m["rows"][r0 + i * L:r0 + (i + 1) * L]).decode() for i in range(nr)]
    v0, v1 = int(m["var_off"][b]), int(m["var_off"][b + 1])
    cols = m["var_col"][v0:v1]
    kinds = m["var_kind"][v0:v1]
    c0 = int(m["cls_off"][b])
    cls = m["cls"][c0:c0 + (v1 - v0) * nr].reshape(v1 - v0, nr) if nr else np.zeros((0, 0), np.uint16)
    i0, i1 = int(m["ilen_off"][b]), int(m["ilen_off"][b + 1])
    return dict(status=int(m["status"][b]), rows=rows,
                snp_pos=[int(c) for c, k in zip(cols, kinds) if k == 0],
                indel_pos=[int(c) for c, k in zip(cols, kinds) if k == 1],
                indel_len=[int(x) for x in m["ilen"][i0:i1]],
                partition={int(c): [int(x) for x in cls[j]] for j, c in enumerate(cols)})


def flatten_seqs(seqs):
    """list[str|bytes] -> (bases uint8 array, seq_off uint64 array)."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return bases, off


def flatten_bubbles(bubbles):
    """list[list[str]] -> (bases, seq_off, bubble_off)."""
    flat = [s for b in bubbles for s in b]
    bases, off = flatten_seqs(flat)
    boff = np.zeros(len(bubbles) + 1, dtype=np.uint32)
    if bubbles:
        boff[1:] = np.cumsum([len(b) for b in bubbles], dtype=np.uint32)
    return bases, off, boff


def kmer_counts_total(seq_off: np.ndarray, k: int) -> int:
    ln = (seq_off[1:] - seq_off[:-1]).astype(np.int64)
    return int(np.maximum(ln - k + 1, 0).sum())


class Checker:
    def __init__(self, which: str):
        assert which in ("oracle", "ref")
        self.which = which
        path = os.path.join(HERE, "libpforacle.so") if which == "oracle" else os.path.join(HERE, "_ref", "libpfref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        self.lib = C.CDLL(path)
        p = "pforc_" if which == "oracle" else "pfref_"
        L = self.lib
        self._open = getattr(L, p + "kmc_open"); self._open.restype = C.c_void_p; self._open.argtypes = [C.c_char_p]
        self._close = getattr(L, p + "kmc_close"); self._close.restype = None; self._close.argtypes = [C.c_void_p]
        self._info = getattr(L, p + "kmc_info"); self._info.restype = C.c_int
        self._info.argtypes = [C.c_void_p, C.POINTER(KmcInfo)]
        self._setmin = getattr(L, p + "kmc_set_min_count"); self._setmin.argtypes = [C.c_void_p, C.c_uint32]
        self._setmax = getattr(L, p + "kmc_set_max_count"); self._setmax.argtypes = [C.c_void_p, C.c_uint32]
        self._counts = getattr(L, p + "kmc_counts"); self._counts.restype = C.c_int
        self._counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]
        self._cov = getattr(L, p + "kmc_cov"); self._cov.restype = C.c_int
        self._cov.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                              C.c_int, C.c_void_p]
        self._align = getattr(L, p + "align"); self._align.restype = C.c_void_p
        self._align.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_uint32, C.c_int, C.POINTER(MsaBatch)]
        self._msa_free = getattr(L, p + "msa_free"); self._msa_free.restype = None
        self._msa_free.argtypes = [C.c_void_p]

    # ---- KMC ----
    def kmc_open(self, prefix: str):
        h = self._open(prefix.encode())
        if not h:
            raise IOError("cannot open KMC database " + prefix)
        return h

    def kmc_close(self, h):
        self._close(h)

    def kmc_info(self, h) -> dict:
        i = KmcInfo()
        if self._info(h, C.byref(i)) != 0:
            raise RuntimeError("kmc_info failed")
        return i.as_dict()

    def kmc_set_min_count(self, h, x):
        self._setmin(h, x)

    def kmc_set_max_count(self, h, x):
        self._setmax(h, x)

    def kmc_counts(self, h, bases, seq_off, k, mode=0, use_read_api=True, n_threads=1):
        n = kmer_counts_total(seq_off, k)
        counts = np.zeros(max(n, 1), dtype=np.uint32)
        found = np.zeros(max(n, 1), dtype=np.uint8)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        rc = self._counts(h, bases.ctypes.data, seq_off.ctypes.data, len(seq_off) - 1, mode, int(use_read_api),
                          n_threads, counts.ctypes.data, found.ctypes.data)
        if rc != 0:
            raise RuntimeError("kmc_counts failed")
        return counts[:n], found[:n]

    def kmc_cov(self, h, bases, seq_off, mode=1, low=0, up=0xFFFFFFFF, n_threads=1):
        n = len(seq_off) - 1
        out = np.zeros(max(n, 1), dtype=COV_DTYPE)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        rc = self._cov(h, bases.ctypes.data, seq_off.ctypes.data, n, mode, low, up, n_threads, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("kmc_cov failed")
        return out[:n]

    # ---- SeqAlign ----
    def align(self, bases, seq_off, bubble_off, M=2.0, D=-1.0, G=-3.0, n_threads=1) -> dict:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        bubble_off = np.ascontiguousarray(bubble_off, dtype=np.uint32)
        mb = MsaBatch()
        h = self._align(M, D, G, bases.ctypes.data, seq_off.ctypes.data, bubble_off.ctypes.data,
                        len(bubble_off) - 1, n_threads, C.byref(mb))
        try:
            return msa_to_numpy(mb)
        finally:
            self._msa_free(h)

    def align_bubbles(self, bubbles, M=2.0, D=-1.0, G=-3.0, n_threads=1) -> dict:
        return self.align(*flatten_bubbles(bubbles), M=M, D=D, G=G, n_threads=n_threads)
