"""Synthetic KMC database writer (.kmc_pre / .kmc_suf), both on-disk dialects.

The `kmc` counter itself is not part of the reference tree (only its reader API is vendored), so the
databases used by tests and bench are generated here in exactly the layout the reference reader
parses: KMC/kmc_api/kmc_file.cpp:185-302 (ReadParamsFrom_prefix_file_buf), :330-366 (CheckKmer),
:1383-1462 (BinarySearch) and mmer.h:34-87 (signature normalisation).  SURVEY.md Appendix A is the
byte-level description.  This is data tooling, not the hot path.

k-mers are handled as numpy uint64 values, 2 bits per symbol, first symbol most significant
(A=0 C=1 G=2 T=3), i.e. value = sum(code[i] << 2*(k-1-i)); k <= 32.
"""
from __future__ import annotations

import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


def encode_bases(seq) -> np.ndarray:
    """bytes/str/uint8 array of characters -> uint8 codes (255 for non-ACGT)."""
    if isinstance(seq, str):
        seq = seq.encode()
    a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)
    return _CODE[a]


def kmers_of(codes: np.ndarray, k: int) -> np.ndarray:
    """All k-mer values of a code array (no N handling: caller guarantees codes < 4)."""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    v = np.zeros(n, dtype=np.uint64)
    c = codes.astype(np.uint64)
    for i in range(k):
        v = (v << np.uint64(2)) | c[i:i + n]
    return v


def revcomp(v: np.ndarray, k: int) -> np.ndarray:
    """Reverse complement of packed k-mers."""
    v = np.asarray(v, dtype=np.uint64)
    x = ~v
    # reverse 2-bit groups inside a 64-bit word
    x = ((x >> np.uint64(2)) & np.uint64(0x3333333333333333)) | ((x & np.uint64(0x3333333333333333)) << np.uint64(2))
    x = ((x >> np.uint64(4)) & np.uint64(0x0F0F0F0F0F0F0F0F)) | ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4))
    x = x.byteswap()
    return x >> np.uint64(64 - 2 * k)


def canonical(v: np.ndarray, k: int) -> np.ndarray:
    return np.minimum(v, revcomp(v, k))


def _is_allowed(m: np.ndarray, length: int) -> np.ndarray:
    """mmer.h:34-57, vectorised over all m-mers of one length."""
    m = m.astype(np.uint32)
    ok = np.ones(m.shape, dtype=bool)
    ok &= (m & 0x3F) != 0x3F          # TTT suffix
    ok &= (m & 0x3F) != 0x3B          # TGT suffix
    ok &= (m & 0x3C) != 0x3C          # TG* suffix
    x = m.copy()
    for _ in range(length - 3):
        ok &= (x & 0xF) != 0          # AA inside
        x = x >> 2
    ok &= x != 0                      # AAA prefix
    ok &= x != 0x04                   # ACA prefix
    ok &= (x & 0xF) != 0              # *AA prefix
    return ok


_NORM_CACHE: dict[int, np.ndarray] = {}


def norm_table(length: int) -> np.ndarray:
    """mmer.h:77-87: norm[x] = min(allowed(x) ? x : 4^m, allowed(rc(x)) ? rc(x) : 4^m)."""
    if length in _NORM_CACHE:
        return _NORM_CACHE[length]
    special = 1 << (2 * length)
    i = np.arange(special, dtype=np.uint64)
    rev = revcomp(i, length).astype(np.uint32)
    i32 = i.astype(np.uint32)
    sv = np.where(_is_allowed(i32, length), i32, np.uint32(special))
    rv = np.where(_is_allowed(rev, length), rev, np.uint32(special))
    t = np.minimum(sv, rv).astype(np.uint32)
    _NORM_CACHE[length] = t
    return t


def signatures(v: np.ndarray, k: int, sig_len: int) -> np.ndarray:
    """kmer_api.h:653-672: min over the k-m+1 m-mers of norm[m-mer]."""
    norm = norm_table(sig_len)
    mask = np.uint64((1 << (2 * sig_len)) - 1)
    best = None
    for i in range(k - sig_len + 1):
        m = (v >> np.uint64(2 * (k - sig_len - i))) & mask
        val = norm[m.astype(np.int64)]
        best = val if best is None else np.minimum(best, val)
    return best


def default_signature_map(sig_len: int, n_bins: int) -> np.ndarray:
    """signature (incl. the sentinel 4^m) -> bin id.  Any map is legal for the reader; this one spreads."""
    n = (1 << (2 * sig_len)) + 1
    s = np.arange(n, dtype=np.uint64)
    return (((s * np.uint64(2654435761)) >> np.uint64(7)) % np.uint64(n_bins)).astype(np.uint32)


def write_kmc_db(prefix: str, kmers: np.ndarray, counts: np.ndarray, k: int, *, version: int = 0x200,
                 lut_prefix_len: int = 5, counter_size: int = 2, sig_len: int = 9, n_bins: int = 64,
                 min_count: int = 1, max_count: int = 10000, both_strands: bool = True,
                 signature_map: np.ndarray | None = None) -> dict:
    """Write <prefix>.kmc_pre and <prefix>.kmc_suf.  `kmers` must be unique (canonical if both_strands)."""
    kmers = np.asarray(kmers, dtype=np.uint64)
    counts = np.asarray(counts, dtype=np.uint64)
    assert kmers.shape == counts.shape
    p = lut_prefix_len
    assert 0 < k <= 32 and (k - p) % 4 == 0 and p >= 1
    S = (k - p) // 4
    C = counter_size
    N = len(kmers)
    pref = (kmers >> np.uint64(2 * (k - p))).astype(np.int64)
    if version == 0x200:
        if signature_map is None:
            signature_map = default_signature_map(sig_len, n_bins)
        bins = signature_map[signatures(kmers, k, sig_len).astype(np.int64)].astype(np.int64)
    elif version == 0:
        bins = np.zeros(N, dtype=np.int64)
        n_bins = 1
    else:
        raise ValueError("version must be 0 (KMC1) or 0x200 (KMC2)")
    order = np.lexsort((kmers, bins))
    kmers, counts, pref, bins = kmers[order], counts[order], pref[order], bins[order]
    # suffix records: S suffix bytes MSB first, then C counter bytes little-endian
    rec = np.zeros((N, S + C), dtype=np.uint8)
    suf = kmers & np.uint64((1 << (2 * (k - p))) - 1) if k - p < 32 else kmers
    for j in range(S):
        rec[:, j] = ((suf >> np.uint64(8 * (S - 1 - j))) & np.uint64(0xFF)).astype(np.uint8)
    for b in range(C):
        rec[:, S + b] = ((counts >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8)
    with open(prefix + ".kmc_suf", "wb") as f:
        f.write(b"KMCS")
        f.write(rec.tobytes())
        f.write(b"KMCS")
    lut_n = n_bins * (1 << (2 * p))
    key = bins * (1 << (2 * p)) + pref
    lut = np.searchsorted(key, np.arange(lut_n, dtype=np.int64), side="left").astype(np.uint64)
    with open(prefix + ".kmc_pre", "wb") as f:
        f.write(b"KMCP")
        f.write(lut.tobytes())
        if version == 0x200:
            f.write(np.uint64(N).tobytes())                       # guard (reader overwrites with N+1)
            f.write(np.asarray(signature_map, dtype=np.uint32).tobytes())
            hdr = np.array([k, 0, C, p, sig_len, min_count, min(max_count, 0xFFFFFFFF)], dtype=np.uint32).tobytes()
            hdr += np.uint64(N).tobytes() + bytes([0 if both_strands else 1])
            hdr += b"\0" * (60 - len(hdr)) + np.uint32(0x200).tobytes()
            f.write(hdr)
            f.write(np.uint32(len(hdr)).tobytes())
        else:
            w = [k | (0 << 32), C | (p << 32), min_count | ((max_count & 0xFFFFFFFF) << 32), N,
                 (0 if both_strands else 1) | ((max_count >> 32) << 32), 0, 0]
            hdr = np.array(w, dtype=np.uint64).tobytes()
            f.write(hdr)
            f.write(np.uint32(len(hdr)).tobytes())
        f.write(b"KMCP")
    return dict(k=k, p=p, S=S, C=C, N=N, version=version, n_bins=n_bins, sig_len=sig_len)


def count_canonical_kmers(seqs, k: int, cap: int = 10000):
    """Canonical k-mer multiset of a list of ACGT sequences -> (sorted unique kmers, counts capped)."""
    parts = []
    for s in seqs:
        codes = encode_bases(s)
        assert (codes < 4).all(), "count_canonical_kmers expects ACGT only"
        parts.append(canonical(kmers_of(codes, k), k))
    allk = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint64)
    u, c = np.unique(allk, return_counts=True)
    return u, np.minimum(c, cap).astype(np.uint64)
