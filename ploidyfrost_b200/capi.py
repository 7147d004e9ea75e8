"""ctypes bindings of libpfgpu.so (include/pf_gpu.h).  Thin: numpy arrays in, numpy arrays out.

There is deliberately no CPU path here: if the library is missing or no GPU is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libpfgpu.so")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)

LOOKUP_CANONICAL, LOOKUP_FWD_THEN_RC, LOOKUP_FWD = 0, 1, 2

# every symbol include/pf_gpu.h declares (tests check that the library exports all of them)
EXPORTS = ["pf_init", "pf_shutdown", "pf_last_error", "pf_version", "pf_launch_count", "pf_sync", "pf_kmc_open",
           "pf_kmc_close", "pf_kmc_info", "pf_kmc_set_min_count", "pf_kmc_set_max_count", "pf_kmc_reset_min_max",
           "pf_kmc_device_bytes", "pf_kmc_open_ex", "pf_kmc_index_kind", "pf_kmc_build_status", "pf_kmc_open_part", "pf_kmc_open_part_ex", "pf_kmc_export_ipc", "pf_kmc_attach_peers", "pf_kmc_local_kmers", "pf_kmc_route_dev", "pf_kmc_lookup_keys_dev",
           "pf_kmc_scatter_dev", "pf_kmc_counts", "pf_kmc_cov", "pf_kmc_cov_async", "pf_kmc_wait", "pf_site_cov", "pf_site_cov_dev", "pf_kmc_lookup_dev", "pf_window_offsets", "pf_align",
           "pf_align_dev", "pf_align_last_tier_counts", "pf_align_last_retry_count", "pf_align_last_heavy_queued", "pf_align_last_cells", "pf_bench_random_gather", "pf_bench_int32", "pf_kmc_share", "pf_bench_gather_sweep", "pf_site_kmers", "pf_lookup_partition", "pf_lookup_partition_sms", "pf_align_staged"]


class SiteBatch(C.Structure):
    _fields_ = [("n_bubbles", C.c_uint32), ("reserved", C.c_uint32), ("site_off", C.POINTER(C.c_uint64)),
                ("status", C.POINTER(C.c_uint8)), ("n_class", C.POINTER(C.c_uint8)), ("cov_off", C.POINTER(C.c_uint64)),
                ("cov", C.POINTER(C.c_uint64))]


SITE_OK, SITE_DROPPED, SITE_MISSING, SITE_UNDEFINED, SITE_SKIPPED = 0, 1, 2, 3, 4


class SiteKmers(C.Structure):
    _fields_ = [("n_bubbles", C.c_uint32), ("reserved", C.c_uint32), ("site_off", C.POINTER(C.c_uint64)),
                ("key_off", C.POINTER(C.c_uint64)), ("keys", C.POINTER(C.c_uint64)), ("status", C.POINTER(C.c_uint8))]


class KmcInfo(C.Structure):
    _fields_ = [("kmer_length", C.c_uint32), ("mode", C.c_uint32), ("counter_size", C.c_uint32),
                ("lut_prefix_length", C.c_uint32), ("signature_len", C.c_uint32), ("min_count", C.c_uint32),
                ("max_count", C.c_uint64), ("total_kmers", C.c_uint64), ("both_strands", C.c_uint32),
                ("kmc_version", C.c_uint32), ("n_bins", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


COV_DTYPE = np.dtype([("sum", "<u8"), ("min", "<u4"), ("n_kmers", "<u4"), ("first_missing", "<i4"),
                      ("first_outside", "<i4")])


class MsaBatch(C.Structure):
    _fields_ = [("n_bubbles", C.c_uint32), ("reserved", C.c_uint32), ("status", i32p), ("n_rows", u32p),
                ("aln_len", u32p), ("rows_off", u64p), ("rows", C.POINTER(C.c_char)), ("var_off", u64p),
                ("var_col", u32p), ("var_kind", u8p), ("cls_off", u64p), ("cls", u16p), ("ilen_off", u64p),
                ("ilen", u32p)]


class PfError(RuntimeError):
    pass


_lib = None


def load():
    """Loads libpfgpu.so (building it is __graft_entry__.build()'s / ploidyfrost_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PfError(f"{LIB_PATH} not built: run `python -m ploidyfrost_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.pf_last_error.restype = C.c_char_p
    L.pf_version.restype = C.c_char_p
    L.pf_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.pf_shutdown.argtypes = [C.c_void_p]
    L.pf_shutdown.restype = None
    L.pf_launch_count.argtypes = [C.c_void_p]
    L.pf_launch_count.restype = C.c_uint64
    L.pf_sync.argtypes = [C.c_void_p]
    L.pf_kmc_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.pf_kmc_open_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.pf_kmc_index_kind.argtypes = [C.c_void_p]
    L.pf_kmc_build_status.argtypes = [C.c_void_p]
    L.pf_kmc_build_status.restype = C.c_uint32
    L.pf_site_cov.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(SiteBatch)]
    L.pf_site_cov_dev.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pf_kmc_open_part.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
    L.pf_kmc_open_part_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
    L.pf_kmc_export_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.pf_kmc_attach_peers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.pf_kmc_local_kmers.argtypes = [C.c_void_p]
    L.pf_kmc_local_kmers.restype = C.c_uint64
    L.pf_kmc_route_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pf_kmc_lookup_keys_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pf_kmc_scatter_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                     C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pf_kmc_close.argtypes = [C.c_void_p]
    L.pf_kmc_info.argtypes = [C.c_void_p, C.POINTER(KmcInfo)]
    L.pf_kmc_set_min_count.argtypes = [C.c_void_p, C.c_uint32]
    L.pf_kmc_set_max_count.argtypes = [C.c_void_p, C.c_uint32]
    L.pf_kmc_reset_min_max.argtypes = [C.c_void_p]
    L.pf_kmc_device_bytes.argtypes = [C.c_void_p]
    L.pf_kmc_device_bytes.restype = C.c_uint64
    L.pf_kmc_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    L.pf_kmc_cov.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    L.pf_kmc_cov_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    L.pf_kmc_wait.argtypes = [C.c_void_p]
    L.pf_kmc_lookup_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                    C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pf_window_offsets.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.pf_window_offsets.restype = C.c_uint64
    L.pf_align.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                           C.POINTER(MsaBatch)]
    L.pf_align_staged.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.POINTER(MsaBatch)]
    L.pf_align_dev.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_uint64, C.c_void_p,
                               C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(MsaBatch), C.c_void_p]
    L.pf_align_last_retry_count.argtypes = [C.c_void_p]
    L.pf_align_last_retry_count.restype = C.c_uint32
    L.pf_align_last_heavy_queued.argtypes = [C.c_void_p]
    L.pf_align_last_heavy_queued.restype = C.c_uint32
    L.pf_align_last_tier_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.pf_align_last_cells.argtypes = [C.c_void_p]
    L.pf_align_last_cells.restype = C.c_uint64
    L.pf_bench_random_gather.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_double)]
    L.pf_bench_int32.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.pf_bench_gather_sweep.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    L.pf_kmc_share.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.pf_lookup_partition.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.pf_lookup_partition_sms.argtypes = [C.c_void_p]
    L.pf_lookup_partition_sms.restype = C.c_uint32
    L.pf_site_kmers.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(SiteKmers)]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise PfError(f"{what} failed ({rc}): {load().pf_last_error().decode()}")


def window_offsets(seq_off: np.ndarray, k: int) -> np.ndarray:
    ln = (seq_off[1:] - seq_off[:-1]).astype(np.int64)
    out = np.zeros(len(seq_off), dtype=np.uint64)
    out[1:] = np.cumsum(np.maximum(ln - k + 1, 0)).astype(np.uint64)
    return out


def _np_from(ptr, n, dtype, copy=True):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = n * np.dtype(dtype).itemsize
    a = np.ctypeslib.as_array(C.cast(ptr, u8p), shape=(nbytes,)).view(dtype)
    return a.copy() if copy else a


def msa_to_numpy(mb: MsaBatch, copy=True) -> dict:
    """pf_msa_batch_t -> dict of numpy arrays.  copy=False returns views of the context-owned pinned result
    buffers (valid until the next pf_align* call on the same context, exactly as the C ABI states)."""
    n = mb.n_bubbles
    get = lambda ptr, cnt, dt: _np_from(ptr, cnt, dt, copy)
    out = {"n_bubbles": n, "status": get(mb.status, n, np.int32), "n_rows": get(mb.n_rows, n, np.uint32),
           "aln_len": get(mb.aln_len, n, np.uint32)}
    for name in ("rows", "var", "cls", "ilen"):
        out[name + "_off"] = get(getattr(mb, name + "_off"), n + 1, np.uint64) if n else np.zeros(1, np.uint64)
    tr, tv, tc, ti = (int(out[k + "_off"][-1]) for k in ("rows", "var", "cls", "ilen"))
    out["rows"] = get(mb.rows, tr, np.uint8)
    out["var_col"] = get(mb.var_col, tv, np.uint32)
    out["var_kind"] = get(mb.var_kind, tv, np.uint8)
    out["cls"] = get(mb.cls, tc, np.uint16)
    out["ilen"] = get(mb.ilen, ti, np.uint32)
    return out


class Context:
    """pf_ctx: one GPU.  Raises PfError when no CUDA device is available (there is no CPU path)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        _check(self.lib.pf_init(device, C.byref(h)), "pf_init")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.pf_shutdown(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launches(self) -> int:
        return int(self.lib.pf_launch_count(self.h))

    def sync(self):
        _check(self.lib.pf_sync(self.h), "pf_sync")

    def align(self, bases, seq_off, bubble_off, M=2.0, D=-1.0, G=-3.0, copy=True) -> dict:
        """SeqAlign::SequenceAlignment over a bubble batch (pf_align).  copy=False: zero-copy views of the result
        arena, valid until the next align call on this context."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        bubble_off = np.ascontiguousarray(bubble_off, dtype=np.uint32)
        mb = MsaBatch()
        _check(self.lib.pf_align(self.h, M, D, G, bases.ctypes.data, seq_off.ctypes.data, bubble_off.ctypes.data,
                                 len(bubble_off) - 1, C.byref(mb)), "pf_align")
        return msa_to_numpy(mb, copy=copy)

    def align_staged(self, db, first_seq, bubble_off, max_len, max_rows, M=2.0, D=-1.0, G=-3.0, copy=True) -> dict:
        """pf_align_staged: align the branches the preceding db.cov / db.cov_async call left on the device (sequences first_seq ..)"""
        bubble_off = np.ascontiguousarray(bubble_off, dtype=np.uint32)
        mb = MsaBatch()
        _check(self.lib.pf_align_staged(self.h, db.h, M, D, G, first_seq, bubble_off.ctypes.data, len(bubble_off) - 1, max_len, max_rows,
                                        C.byref(mb)), "pf_align_staged")
        return msa_to_numpy(mb, copy=copy)

    def align_dev(self, d_bases, n_bases, d_seq_off, n_seq, d_bubble_off, n_bubbles, max_len, max_rows, M=2.0, D=-1.0,
                  G=-3.0, stream=None) -> MsaBatch:
        mb = MsaBatch()
        _check(self.lib.pf_align_dev(self.h, M, D, G, d_bases, n_bases, d_seq_off, n_seq, d_bubble_off, n_bubbles,
                                     max_len, max_rows, C.byref(mb), stream), "pf_align_dev")
        return mb

    def site_kmers(self, k, skip=None) -> dict:
        """pf_site_kmers: the site k-mers (2-bit keys) of the last alignment of this context, one per (variable column, row)"""
        sk = SiteKmers()
        sp = None
        if skip is not None:
            sp = np.ascontiguousarray(skip, dtype=np.uint8)
        _check(self.lib.pf_site_kmers(self.h, k, sp.ctypes.data if sp is not None else None, C.byref(sk)), "pf_site_kmers")
        n = sk.n_bubbles
        site_off = np.ctypeslib.as_array(sk.site_off, shape=(n + 1,)).copy()
        key_off = np.ctypeslib.as_array(sk.key_off, shape=(n + 1,)).copy()
        ns, nk = int(site_off[-1]), int(key_off[-1])
        return {"site_off": site_off, "key_off": key_off,
                "keys": np.ctypeslib.as_array(sk.keys, shape=(nk,)).copy() if nk else np.zeros(0, np.uint64),
                "status": np.ctypeslib.as_array(sk.status, shape=(ns,)).copy() if ns else np.zeros(0, np.uint8)}

    @property
    def last_retry_count(self) -> int:
        return int(self.lib.pf_align_last_retry_count(self.h))

    @property
    def last_heavy_queued(self) -> int:
        return int(self.lib.pf_align_last_heavy_queued(self.h))

    @property
    def last_tier_counts(self) -> list:
        a = (C.c_uint32 * 9)()
        _check(self.lib.pf_align_last_tier_counts(self.h, a, 9), "pf_align_last_tier_counts")
        return list(a)

    @property
    def last_cells(self) -> int:
        return int(self.lib.pf_align_last_cells(self.h))

    def bench_random_gather(self, nbytes: int) -> float:
        v = C.c_double()
        _check(self.lib.pf_bench_random_gather(self.h, nbytes, C.byref(v)), "pf_bench_random_gather")
        return v.value

    def lookup_partition(self, n_sm: int):
        """pf_lookup_partition -> (raw CUDA stream handle confined to an SM partition, SMs granted)"""
        st = C.c_void_p()
        _check(self.lib.pf_lookup_partition(self.h, n_sm, C.byref(st)), "pf_lookup_partition")
        return st.value, int(self.lib.pf_lookup_partition_sms(self.h))

    def bench_gather_sweep(self, nbytes: int, width: int, ilp: int, ctas_per_sm: int) -> float:
        v = C.c_double()
        _check(self.lib.pf_bench_gather_sweep(self.h, nbytes, width, ilp, ctas_per_sm, C.byref(v)), "pf_bench_gather_sweep")
        return v.value

    def bench_int32(self) -> float:
        v = C.c_double()
        _check(self.lib.pf_bench_int32(self.h, C.byref(v)), "pf_bench_int32")
        return v.value


INDEX_AUTO, INDEX_VERBATIM, INDEX_HASH = 0, 1, 2


class KmcDb:
    """pf_kmc: HBM-resident KMC index (CKMCFile opened for random access)."""

    def __init__(self, ctx: Context, prefix: str, part: int = 0, n_parts: int = 1, index: str = "auto", share_of: "KmcDb" = None):
        """n_parts > 1: load only partition `part` of the database (pf_kmc_open_part); such an index answers
        route_dev / lookup_keys_dev / scatter_dev, not counts / cov.
        share_of: do not open anything -- a second handle on `share_of`'s index for the context `ctx` (pf_kmc_share)."""
        self.ctx = ctx
        self.lib = ctx.lib
        self.part, self.n_parts = part, n_parts
        h = C.c_void_p()
        if share_of is not None:
            self.part, self.n_parts = share_of.part, share_of.n_parts
            _check(self.lib.pf_kmc_share(share_of.h, ctx.h, C.byref(h)), "pf_kmc_share")
        elif n_parts == 1 and index == "verbatim":
            _check(self.lib.pf_kmc_open_ex(ctx.h, prefix.encode(), INDEX_VERBATIM, C.byref(h)), "pf_kmc_open_ex")
        elif n_parts == 1:
            _check(self.lib.pf_kmc_open(ctx.h, prefix.encode(), C.byref(h)), "pf_kmc_open")
        elif index == "verbatim":
            _check(self.lib.pf_kmc_open_part_ex(ctx.h, prefix.encode(), part, n_parts, INDEX_VERBATIM, C.byref(h)), "pf_kmc_open_part_ex")
        else:
            _check(self.lib.pf_kmc_open_part(ctx.h, prefix.encode(), part, n_parts, C.byref(h)), "pf_kmc_open_part")
        self.h = h
        self.refresh_info()
        self.k = self.info["kmer_length"]

    def site_cov_dev(self, low, up, d_skip=None, stream=None):
        """pf_site_cov_dev: asynchronous, results stay on the device."""
        _check(self.lib.pf_site_cov_dev(self.h, low, up, d_skip, None, stream), "pf_site_cov_dev")

    def site_cov(self, low, up, skip=None, copy=True):
        """pf_site_cov: class coverages of the variable columns of the context's last alignment (lookup phase B).
        copy=False returns views of the handle's pinned result arena (valid until the next pf_site_cov on this handle)."""
        sb = SiteBatch()
        sk = None
        if skip is not None:
            sk = np.ascontiguousarray(skip, dtype=np.uint8)
        _check(self.lib.pf_site_cov(self.h, low, up, sk.ctypes.data if sk is not None else None, C.byref(sb)), "pf_site_cov")
        n = sb.n_bubbles
        fin = (lambda a: a.copy()) if copy else (lambda a: a)
        site_off = np.ctypeslib.as_array(sb.site_off, shape=(n + 1,))
        cov_off = np.ctypeslib.as_array(sb.cov_off, shape=(n + 1,))
        ns, nc = int(site_off[-1]), int(cov_off[-1])
        return {"site_off": fin(site_off), "cov_off": fin(cov_off),
                "status": fin(np.ctypeslib.as_array(sb.status, shape=(ns,))) if ns else np.zeros(0, np.uint8),
                "n_class": fin(np.ctypeslib.as_array(sb.n_class, shape=(ns,))) if ns else np.zeros(0, np.uint8),
                "cov": fin(np.ctypeslib.as_array(sb.cov, shape=(nc,))) if nc else np.zeros(0, np.uint64)}

    @property
    def device_bytes(self) -> int:
        return int(self.lib.pf_kmc_device_bytes(self.h))

    def export_ipc(self) -> np.ndarray:
        blob = np.zeros(128, np.uint8)
        _check(self.lib.pf_kmc_export_ipc(self.h, blob.ctypes.data), "pf_kmc_export_ipc")
        return blob

    def attach_peers(self, blobs: np.ndarray):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8)
        assert blobs.size == 128 * self.n_parts
        _check(self.lib.pf_kmc_attach_peers(self.h, blobs.ctypes.data, self.n_parts), "pf_kmc_attach_peers")

    @property
    def build_status(self) -> int:
        return int(self.lib.pf_kmc_build_status(self.h))

    @property
    def index_kind(self) -> str:
        return {INDEX_VERBATIM: "verbatim", INDEX_HASH: "hash"}.get(self.lib.pf_kmc_index_kind(self.h), "?")

    def close(self):
        if self.h:
            self.lib.pf_kmc_close(self.h)
            self.h = None

    def refresh_info(self):
        i = KmcInfo()
        _check(self.lib.pf_kmc_info(self.h, C.byref(i)), "pf_kmc_info")
        self.info = i.as_dict()
        return self.info

    def set_min_count(self, x):
        _check(self.lib.pf_kmc_set_min_count(self.h, x), "pf_kmc_set_min_count")

    def set_max_count(self, x):
        _check(self.lib.pf_kmc_set_max_count(self.h, x), "pf_kmc_set_max_count")

    def reset_min_max(self):
        _check(self.lib.pf_kmc_reset_min_max(self.h), "pf_kmc_reset_min_max")

    @property
    def device_bytes(self) -> int:
        return int(self.lib.pf_kmc_device_bytes(self.h))

    @property
    def local_kmers(self) -> int:
        return int(self.lib.pf_kmc_local_kmers(self.h))

    def route_dev(self, d_bases, n_bases, d_seq_off, d_win_off, n_seq, n_windows, mode, d_send_keys, d_send_idx, stream=None):
        """-> send_off[n_parts + 1] (numpy uint64): bucket boundaries of the routed keys."""
        off = np.zeros(self.n_parts + 1, dtype=np.uint64)
        _check(self.lib.pf_kmc_route_dev(self.h, d_bases, n_bases, d_seq_off, d_win_off, n_seq, n_windows, mode, d_send_keys,
                                         d_send_idx, off.ctypes.data, stream), "pf_kmc_route_dev")
        return off

    def lookup_keys_dev(self, d_keys, n, d_counts, d_found, stream=None):
        _check(self.lib.pf_kmc_lookup_keys_dev(self.h, d_keys, n, d_counts, d_found, stream), "pf_kmc_lookup_keys_dev")

    def scatter_dev(self, d_send_idx, n_sent, d_reply_counts, d_reply_found, d_win_off, n_seq, n_windows, low, up, d_counts,
                    d_found, d_cov=None, stream=None):
        _check(self.lib.pf_kmc_scatter_dev(self.h, d_send_idx, n_sent, d_reply_counts, d_reply_found, d_win_off, n_seq, n_windows,
                                           low, up, d_counts, d_found, d_cov, stream), "pf_kmc_scatter_dev")

    def counts(self, bases, seq_off, mode=LOOKUP_CANONICAL):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        n = int(window_offsets(seq_off, self.k)[-1]) if len(seq_off) > 1 else 0
        counts = np.zeros(max(n, 1), dtype=np.uint32)
        found = np.zeros(max(n, 1), dtype=np.uint8)
        _check(self.lib.pf_kmc_counts(self.h, bases.ctypes.data, seq_off.ctypes.data, len(seq_off) - 1, mode,
                                      counts.ctypes.data, found.ctypes.data), "pf_kmc_counts")
        return counts[:n], found[:n]

    def cov(self, bases, seq_off, mode=LOOKUP_FWD_THEN_RC, low=0, up=0xFFFFFFFF, out=None):
        """readCov reductions, one record per sequence (pf_kmc_cov).  `out`: optional preallocated COV_DTYPE array
        (e.g. over pinned memory, which makes the device->host copy a direct DMA)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        n = len(seq_off) - 1
        if out is None:
            out = np.zeros(max(n, 1), dtype=COV_DTYPE)
        assert out.dtype == COV_DTYPE and len(out) >= n and out.flags.c_contiguous
        _check(self.lib.pf_kmc_cov(self.h, bases.ctypes.data, seq_off.ctypes.data, n, mode, low, up, out.ctypes.data),
               "pf_kmc_cov")
        return out[:n]

    def cov_async(self, bases, seq_off, out, mode=LOOKUP_FWD_THEN_RC, low=0, up=0xFFFFFFFF):
        """pf_kmc_cov_async: returns at once; `bases`, `seq_off`, `out` (contiguous, ideally pinned) must live until wait()."""
        assert bases.dtype == np.uint8 and bases.flags.c_contiguous and seq_off.dtype == np.uint64 and seq_off.flags.c_contiguous
        n = len(seq_off) - 1
        assert out.dtype == COV_DTYPE and len(out) >= n and out.flags.c_contiguous
        _check(self.lib.pf_kmc_cov_async(self.h, bases.ctypes.data, seq_off.ctypes.data, n, mode, low, up, out.ctypes.data),
               "pf_kmc_cov_async")
        return out[:n]

    def wait(self):
        _check(self.lib.pf_kmc_wait(self.h), "pf_kmc_wait")

    def lookup_dev(self, d_bases, n_bases, d_seq_off, d_win_off, n_seq, n_windows, mode=LOOKUP_CANONICAL, low=0,
                   up=0xFFFFFFFF, d_counts=None, d_found=None, d_cov=None, stream=None):
        _check(self.lib.pf_kmc_lookup_dev(self.h, d_bases, n_bases, d_seq_off, d_win_off, n_seq, n_windows, mode, low, up,
                                          d_counts, d_found, d_cov, stream), "pf_kmc_lookup_dev")
