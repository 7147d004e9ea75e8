"""Synthetic workloads of the BASELINE.json configs: haplotypes + flat bubble batches (C++, pfsynth.cpp) and
the matching KMC database (k-mer sort/count on the GPU with torch -- plumbing -- or numpy for small cases).

Data tooling only; nothing here is on the measured hot path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import kmcdb

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpfsynth.so")


def build_synth(force=False):
    src = os.path.join(HERE, "pfsynth.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-pthread", "-shared", "-o", LIB, src], check=True)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_synth()
        L = C.CDLL(LIB)
        L.pfs_create.restype = C.c_void_p
        L.pfs_create.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_uint32,
                                 C.c_uint32, C.c_int]
        L.pfs_create_sub.restype = C.c_void_p
        L.pfs_create_sub.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_uint32,
                                     C.c_uint32, C.c_int, C.c_uint64, C.c_double, C.c_uint64]
        L.pfs_make_long_bubbles.restype = C.c_uint64
        L.pfs_make_long_bubbles.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.pfs_destroy.argtypes = [C.c_void_p]
        L.pfs_hap_len.restype = C.c_uint64
        L.pfs_hap_len.argtypes = [C.c_void_p, C.c_uint32]
        L.pfs_hap_ptr.restype = C.c_void_p
        L.pfs_hap_ptr.argtypes = [C.c_void_p, C.c_uint32]
        L.pfs_anc_ptr.restype = C.c_void_p
        L.pfs_anc_ptr.argtypes = [C.c_void_p]
        L.pfs_make_bubbles.restype = C.c_uint64
        L.pfs_make_bubbles.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
        for name in ("pfs_n_seq", "pfs_n_bases", "pfs_n_ent_bases"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("pfs_bases", "pfs_seq_off", "pfs_bubble_off", "pfs_bubble_type", "pfs_ent_size", "pfs_exit_size",
                     "pfs_ent_bases", "pfs_ent_off"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype).copy()


class BubbleBatch:
    """Flat bubble batch (include/pf_types.h conventions) + the entrance unitigs of the same bubbles."""

    def __init__(self, bases, seq_off, bubble_off, bubble_type, ent_bases, ent_off, ent_size, exit_size):
        self.bases, self.seq_off, self.bubble_off, self.bubble_type = bases, seq_off, bubble_off, bubble_type
        self.ent_bases, self.ent_off, self.ent_size, self.exit_size = ent_bases, ent_off, ent_size, exit_size

    @property
    def n_bubbles(self):
        return len(self.bubble_off) - 1

    @property
    def n_seq(self):
        return len(self.seq_off) - 1

    def lookup_sequences(self):
        """entrance unitigs followed by all branch sequences, as one sequence batch (lookup phase A)."""
        bases = np.concatenate([self.ent_bases, self.bases])
        off = np.concatenate([self.ent_off, self.seq_off[1:] + self.ent_off[-1]]).astype(np.uint64)
        return bases, off

    @staticmethod
    def concat(parts):
        """one batch out of several (bubbles of different subgenomes / regions), in the order given"""
        parts = [p for p in parts if p.n_bubbles]
        if len(parts) == 1:
            return parts[0]

        def cat_off(arrs, dtype):
            out, base = [np.zeros(1, dtype)], 0
            for a in arrs:
                out.append((a[1:].astype(np.uint64) + np.uint64(base)).astype(dtype))
                base += int(a[-1])
            return np.concatenate(out)
        return BubbleBatch(np.concatenate([p.bases for p in parts]), cat_off([p.seq_off for p in parts], np.uint64),
                           cat_off([p.bubble_off for p in parts], np.uint32), np.concatenate([p.bubble_type for p in parts]),
                           np.concatenate([p.ent_bases for p in parts]), cat_off([p.ent_off for p in parts], np.uint64),
                           np.concatenate([p.ent_size for p in parts]), np.concatenate([p.exit_size for p in parts]))

    def slice(self, b0, b1):
        s0, s1 = int(self.bubble_off[b0]), int(self.bubble_off[b1])
        c0, c1 = int(self.seq_off[s0]), int(self.seq_off[s1])
        e0, e1 = int(self.ent_off[b0]), int(self.ent_off[b1])
        return BubbleBatch(self.bases[c0:c1], self.seq_off[s0:s1 + 1] - self.seq_off[s0],
                           self.bubble_off[b0:b1 + 1] - self.bubble_off[b0], self.bubble_type[b0:b1],
                           self.ent_bases[e0:e1], self.ent_off[b0:b1 + 1] - self.ent_off[b0], self.ent_size[b0:b1],
                           self.exit_size[b0:b1])

    def stats(self):
        nr = np.diff(self.bubble_off)
        ln = np.diff(self.seq_off)
        return dict(n_bubbles=int(self.n_bubbles), n_seq=int(self.n_seq), strict=int(self.bubble_type.sum()),
                    rows_hist={int(k): int(v) for k, v in zip(*np.unique(nr, return_counts=True))},
                    len_p50=int(np.percentile(ln, 50)) if len(ln) else 0, len_p99=int(np.percentile(ln, 99)) if len(ln) else 0,
                    len_max=int(ln.max()) if len(ln) else 0,
                    dp_cells=int(self.dp_cells()))

    def dp_cells(self):
        """(m+1)(n+1) cells of the first pairwise fill of every bubble plus one fill per further row (lower bound)."""
        ln = np.diff(self.seq_off).astype(np.int64)
        total = 0
        bo = self.bubble_off.astype(np.int64)
        first = ln[bo[:-1]]
        nr = np.diff(bo)
        # rows 1..n-1 are each aligned against row 0 (or its gapped profile): >= (len0+1)*(len_i+1) cells
        rep_first = np.repeat(first, nr)
        cells = (rep_first + 1) * (ln + 1)
        cells[bo[:-1]] = 0
        total = int(cells.sum())
        return total


class Workload:
    def __init__(self, seed, genome_len, n_hap, p_snp=0.01, p_indel=0.001, p_long=0.0, long_min=100, long_max=5000,
                 n_threads=8, root_seed=0, divergence=0.0, offset=0):
        """root_seed != 0: a subgenome -- the ancestor is the root genome of `root_seed` with `divergence` substitutions.
        offset: generate only the stretch [offset, offset + genome_len) of the whole genome (same bases, same variants away
        from the first / last ~6 kbp of the stretch)."""
        self.L = _load()
        self.h = self.L.pfs_create_sub(seed, genome_len, n_hap, p_snp, p_indel, p_long, long_min, long_max, n_threads,
                                       root_seed, divergence, offset)
        self.n_hap = n_hap
        self.genome_len = genome_len

    def close(self):
        if self.h:
            self.L.pfs_destroy(self.h)
            self.h = None

    def haplotype(self, i) -> np.ndarray:
        return _arr(self.L.pfs_hap_ptr(self.h, i), self.L.pfs_hap_len(self.h, i), np.uint8)

    def ancestor(self) -> np.ndarray:
        return _arr(self.L.pfs_anc_ptr(self.h), self.genome_len, np.uint8)

    def long_bubbles(self, seed, k, n_bubbles, min_len=50, max_len=5000, max_rows=4) -> BubbleBatch:
        """BASELINE configs[4]: branch lengths log-uniform in [min_len, max_len], >= 1 long indel per bubble"""
        self.L.pfs_make_long_bubbles(self.h, seed, k, n_bubbles, min_len, max_len, max_rows)
        return self._batch(n_bubbles)

    def bubbles(self, k, r0, r1, max_bubbles=1 << 62) -> BubbleBatch:
        nb = self.L.pfs_make_bubbles(self.h, k, r0, r1, max_bubbles)
        return self._batch(nb)

    def _batch(self, nb) -> BubbleBatch:
        ns = self.L.pfs_n_seq(self.h)
        h = self.h
        return BubbleBatch(_arr(self.L.pfs_bases(h), self.L.pfs_n_bases(h), np.uint8),
                           _arr(self.L.pfs_seq_off(h), ns + 1, np.uint64), _arr(self.L.pfs_bubble_off(h), nb + 1, np.uint32),
                           _arr(self.L.pfs_bubble_type(h), nb, np.uint8),
                           _arr(self.L.pfs_ent_bases(h), self.L.pfs_n_ent_bases(h), np.uint8),
                           _arr(self.L.pfs_ent_off(h), nb + 1, np.uint64), _arr(self.L.pfs_ent_size(h), nb, np.uint32),
                           _arr(self.L.pfs_exit_size(h), nb, np.uint32))


def write_db_numpy(prefix, haps, k, lam_per_copy, seed, **kw):
    """Small cases: canonical k-mers of all haplotypes, count ~ Poisson(multiplicity * lam) clamped to [1, 10000]."""
    parts = [kmcdb.canonical(kmcdb.kmers_of(kmcdb.encode_bases(h), k), k) for h in haps]
    u, mult = np.unique(np.concatenate(parts), return_counts=True)
    rng = np.random.default_rng(seed)
    cnt = np.clip(rng.poisson(mult * lam_per_copy), 1, 10000).astype(np.uint64)
    info = kmcdb.write_kmc_db(prefix, u, cnt, k, **kw)
    return info, u, cnt


def _canonical_kmers_torch(h, k, dev, code):
    import torch
    c = code[torch.from_numpy(h).to(dev).long()]
    n = c.numel() - k + 1
    fwd = torch.zeros(n, dtype=torch.int64, device=dev)
    rc = torch.zeros(n, dtype=torch.int64, device=dev)
    for i in range(k):
        fwd = (fwd << 2) | c[i:i + n]
        rc = rc | ((3 - c[i:i + n]) << (2 * i))
    return torch.minimum(fwd, rc)


def write_db_torch(prefix, haps, k, lam_per_copy, seed, device="cuda", version=0x200, lut_prefix_len=9, sig_len=9,
                   n_bins=512, counter_size=2, min_count=1, max_count=10000):
    """Large cases: the same database built with torch on the GPU (sort/unique of the packed canonical k-mers).
    `haps` is a list of haplotypes, or a list of GROUPS of haplotypes (subgenomes): every group is de-duplicated on its own
    and the groups are merged afterwards, so that no single sort sees more than one group's k-mers (configs[2]: 3 groups of
    2 x 333 Mbp, 1.2 G distinct k-mers)."""
    import torch
    dev = torch.device(device)
    code = torch.full((256,), 0, dtype=torch.int64, device=dev)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    groups = haps if isinstance(haps[0], (list, tuple)) else [haps]
    us, ms = [], []
    for grp in groups:
        allk = torch.cat([_canonical_kmers_torch(h, k, dev, code) for h in grp])
        u, mult = torch.unique(allk, sorted=True, return_counts=True)
        del allk
        us.append(u)
        ms.append(mult)
    if len(us) == 1:
        u, mult = us[0], ms[0]
    else:   # merge the groups: k-mers shared between subgenomes add their multiplicities
        keys, mm = torch.cat(us), torch.cat(ms)
        del us, ms
        keys, order = torch.sort(keys)
        mm = mm[order]
        del order
        first = torch.ones(keys.numel(), dtype=torch.bool, device=dev)
        first[1:] = keys[1:] != keys[:-1]
        seg = torch.cumsum(first, 0) - 1
        u = keys[first]
        del keys
        mult = torch.zeros(u.numel(), dtype=torch.int64, device=dev).index_add_(0, seg, mm)
        del seg, mm, first
    us = ms = None
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    cnt = torch.poisson(mult.double() * lam_per_copy, generator=g).clamp_(1, 10000).long()
    del mult
    N = u.numel()
    p = lut_prefix_len
    S = (k - p) // 4
    assert (k - p) % 4 == 0 and k <= 31
    pref = u >> (2 * (k - p))
    if version == 0x200:
        norm = torch.from_numpy(kmcdb.norm_table(sig_len).astype(np.int64)).to(dev)
        smask = (1 << (2 * sig_len)) - 1
        sig = None
        for i in range(k - sig_len + 1):
            v = norm[(u >> (2 * (k - sig_len - i))) & smask]
            sig = v if sig is None else torch.minimum(sig, v)
        del v
        smap_np = kmcdb.default_signature_map(sig_len, n_bins)
        bins = torch.from_numpy(smap_np.astype(np.int64)).to(dev)[sig]
        del sig
    else:
        n_bins = 1
        smap_np = None
        bins = torch.zeros_like(u)
    slot = bins * (1 << (2 * p)) + pref
    del bins, pref
    order = torch.argsort((slot << (2 * (k - p))) | (u & ((1 << (2 * (k - p))) - 1)))   # (bin, prefix, suffix)
    u, cnt, slot = u[order], cnt[order], slot[order]
    del order
    lut = torch.searchsorted(slot, torch.arange(n_bins * (1 << (2 * p)), device=dev, dtype=torch.int64))
    del slot
    lut_np = lut.cpu().numpy().astype(np.uint64)
    del lut
    suf = u & ((1 << (2 * (k - p))) - 1)
    del u
    with open(prefix + ".kmc_suf", "wb") as f:   # records in chunks: the host never holds more than one chunk
        f.write(b"KMCS")
        step = 1 << 26
        for a in range(0, N, step):
            b = min(N, a + step)
            rec = torch.empty((b - a, S + counter_size), dtype=torch.uint8, device=dev)
            for j in range(S):
                rec[:, j] = ((suf[a:b] >> (8 * (S - 1 - j))) & 0xFF).to(torch.uint8)
            for q in range(counter_size):
                rec[:, S + q] = ((cnt[a:b] >> (8 * q)) & 0xFF).to(torch.uint8)
            rec.cpu().numpy().tofile(f)
            del rec
        f.write(b"KMCS")
    del suf, cnt
    with open(prefix + ".kmc_pre", "wb") as f:
        f.write(b"KMCP")
        lut_np.tofile(f)
        if version == 0x200:
            f.write(np.uint64(N).tobytes())
            f.write(np.asarray(smap_np, dtype=np.uint32).tobytes())
            hdr = np.array([k, 0, counter_size, p, sig_len, min_count, max_count], dtype=np.uint32).tobytes()
            hdr += np.uint64(N).tobytes() + bytes([0])
            hdr += b"\0" * (60 - len(hdr)) + np.uint32(0x200).tobytes()
        else:
            hdr = np.array([k, counter_size | (p << 32), min_count | (max_count << 32), N, 0, 0, 0], dtype=np.uint64).tobytes()
        f.write(hdr)
        f.write(np.uint32(len(hdr)).tobytes())
        f.write(b"KMCP")
    torch.cuda.empty_cache() if dev.type == "cuda" else None
    return dict(k=k, p=p, S=S, C=counter_size, N=int(N), version=version, n_bins=n_bins, sig_len=sig_len)
