"""oracle/parity.py -- TEST INFRASTRUCTURE: field-by-field comparison of the CUDA path's results with a CPU checker's.

Used by tests/ (the large differential tests) and by bench.py's parity block (every timed number carries the parity of the
very batch it was timed on).  Never imported by the product package.

  compare_cov(a, b)               pf_cov_t records (CDBG::readCov reductions, CDBG.cpp:29-120)
  compare_msa(a, b)               every field of pf_msa_batch_t, per bubble (SeqAlign::SequenceAlignment, SeqAlign.cpp:550-640)
  expected_site_cov(...)          lookup phase B (CDBG.cpp:2295-2509) from an alignment result + a per-k-mer lookup function:
                                  vectorised for SNP sites that no indel site precedes (:2469-2472), oracle/caller.py (pure
                                  Python, the pinned restatement) for a bounded number of the other sites
"""
from __future__ import annotations

import numpy as np

from . import caller

SITE_OK, SITE_DROPPED, SITE_MISSING, SITE_UNDEFINED, SITE_SKIPPED = 0, 1, 2, 3, 4


def compare_cov(a: np.ndarray, b: np.ndarray) -> int:
    """number of pf_cov_t records that differ in any field"""
    if len(a) != len(b):
        return max(len(a), len(b))
    bad = np.zeros(len(a), dtype=bool)
    for f in ("sum", "min", "n_kmers", "first_missing", "first_outside"):
        bad |= a[f] != b[f]
    return int(bad.sum())


def _ragged_equal(xa, oa, xb, ob, width=1):
    """per segment: are xa[oa[i]*w:oa[i+1]*w] and xb[ob[i]*w:ob[i+1]*w] equal?  (vectorised)"""
    n = len(oa) - 1
    la, lb = np.diff(oa.astype(np.int64)), np.diff(ob.astype(np.int64))
    ok = la == lb
    if not ok.all():
        # compare only the segments of equal length; the others are mismatches already
        idx = np.flatnonzero(ok)
    else:
        idx = np.arange(n)
    if len(idx) == 0:
        return ok
    ln = la[idx] * width
    tot = int(ln.sum())
    if tot == 0:
        return ok
    seg = np.repeat(np.arange(len(idx)), ln)
    within = np.arange(tot) - np.repeat(np.cumsum(ln) - ln, ln)
    pa = oa.astype(np.int64)[idx][seg] * width + within
    pb = ob.astype(np.int64)[idx][seg] * width + within
    neq = xa[pa] != xb[pb]
    bad_seg = np.zeros(len(idx), dtype=bool)
    np.logical_or.at(bad_seg, seg[neq], True)
    ok[idx[bad_seg]] = False
    return ok


def compare_msa(a: dict, b: dict) -> np.ndarray:
    """-> bool array, True where bubble i differs in ANY field of pf_msa_batch_t (status, n_rows, aln_len, the aligned rows, the
    variable columns with their kinds, the class ids, the indel lengths)."""
    n = int(a["n_bubbles"])
    if n != int(b["n_bubbles"]):
        return np.ones(max(n, int(b["n_bubbles"])), dtype=bool)
    ok = (a["status"] == b["status"]) & (a["n_rows"] == b["n_rows"]) & (a["aln_len"] == b["aln_len"])
    ok &= _ragged_equal(a["rows"], a["rows_off"], b["rows"], b["rows_off"])
    ok &= _ragged_equal(a["var_col"], a["var_off"], b["var_col"], b["var_off"])
    ok &= _ragged_equal(a["var_kind"], a["var_off"], b["var_kind"], b["var_off"])
    ok &= _ragged_equal(a["cls"], a["cls_off"], b["cls"], b["cls_off"])
    ok &= _ragged_equal(a["ilen"], a["ilen_off"], b["ilen"], b["ilen_off"])
    return ~ok


_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c + 32] = _i


def expected_site_cov(msa: dict, skip, k: int, low: int, up: int, lookup, max_general: int = 4000):
    """What pf_site_cov must return for the variable columns of `msa` (in var_off order).

    lookup(bases_u8_flat, seq_off_u64) -> (counts, found) per k-mer, 'as written, else reverse complement' (CDBG.cpp:38-43).
    Returns (checked bool[n_sites], status uint8[n_sites], n_class uint8[n_sites], cov uint64[cls total]); `cov` has n_rows
    entries per site, the first n_class of them meaningful (the layout of pf_site_batch_t).  Sites outside `checked` (general
    sites beyond max_general) carry no expectation."""
    nb = int(msa["n_bubbles"])
    var_off = msa["var_off"].astype(np.int64)
    cls_off = msa["cls_off"].astype(np.int64)
    nv = np.diff(var_off)
    n_sites = int(var_off[-1])
    b_of = np.repeat(np.arange(nb), nv)
    col = msa["var_col"].astype(np.int64)
    kind = msa["var_kind"]
    nr_b = msa["n_rows"].astype(np.int64)
    L_b = msa["aln_len"].astype(np.int64)
    rows_off = msa["rows_off"].astype(np.int64)
    nr = nr_b[b_of]
    # class entries of site v start at cls_off[b] + (v - var_off[b]) * nr
    v_in_b = np.arange(n_sites) - var_off[b_of]
    cls_at = cls_off[b_of] + v_in_b * nr
    # number of indel sites before each site inside its bubble
    is_ind = (kind == 1).astype(np.int64)
    cum = np.cumsum(is_ind) - is_ind
    n_ind_before = cum - cum[var_off[b_of]] if n_sites else cum
    skipped = np.asarray(skip, dtype=bool)[b_of] if skip is not None else np.zeros(n_sites, dtype=bool)

    status = np.zeros(n_sites, dtype=np.uint8)
    ncls = np.zeros(n_sites, dtype=np.uint8)
    cov = np.zeros(int(cls_off[-1]), dtype=np.uint64)
    checked = np.zeros(n_sites, dtype=bool)
    cls = msa["cls"]
    # n_class of every site = max class id over its rows
    if n_sites:
        site_of_entry = np.repeat(np.arange(n_sites), nr)
        mx = np.zeros(n_sites, dtype=np.int64)
        np.maximum.at(mx, site_of_entry, cls.astype(np.int64))
        ncls = np.minimum(mx, 255).astype(np.uint8)
    status[skipped] = SITE_SKIPPED
    checked[skipped] = True

    simple = (~skipped) & (kind == 0) & (n_ind_before == 0) & (col - k + 1 >= 0) & (nr <= 16)
    idx = np.flatnonzero(simple)
    if len(idx):
        nrs = nr[idx]
        e_site = np.repeat(np.arange(len(idx)), nrs)                     # event -> simple-site ordinal
        e_row = np.arange(len(e_site)) - np.repeat(np.cumsum(nrs) - nrs, nrs)
        sb = b_of[idx][e_site]
        start = rows_off[sb] + e_row * L_b[sb] + col[idx][e_site] - k + 1
        win = msa["rows"][start[:, None] + np.arange(k)[None, :]]
        codes = _CODE[win]
        bad_win = (codes > 3).any(axis=1)
        key = np.zeros(len(win), dtype=np.uint64)
        for j in range(k):
            key = (key << np.uint64(2)) | (codes[:, j] & 3).astype(np.uint64)
        e_cls = cls[(cls_at[idx][e_site] + e_row)].astype(np.int64)
        cnt, fnd = lookup(np.ascontiguousarray(win.reshape(-1)), np.arange(len(win) + 1, dtype=np.uint64) * np.uint64(k))
        cnt = cnt.astype(np.int64)
        fnd = fnd.astype(bool)
        # distinct (site, class, key), in the reference's order: classes ascending, k-mers in std::set (= numeric) order
        order = np.lexsort((key, e_cls, e_site))
        s_site, s_cls, s_key = e_site[order], e_cls[order], key[order]
        first = np.ones(len(order), dtype=bool)
        first[1:] = (s_site[1:] != s_site[:-1]) | (s_cls[1:] != s_cls[:-1]) | (s_key[1:] != s_key[:-1])
        ev = order[first]
        ev_site, ev_cls = e_site[ev], e_cls[ev]
        ev_cnt, ev_fnd = cnt[ev], fnd[ev]
        fail = np.where(~ev_fnd, SITE_MISSING, np.where((ev_cnt > low) & (ev_cnt < up), SITE_OK, SITE_DROPPED))
        pos = np.arange(len(ev))
        big = len(ev) + 1
        first_fail = np.full(len(idx), big, dtype=np.int64)
        np.minimum.at(first_fail, ev_site[fail != 0], pos[fail != 0])
        st = np.zeros(len(idx), dtype=np.uint8)
        hit = first_fail < big
        st[hit] = fail[first_fail[hit]]
        valid = pos < first_fail[ev_site]
        tgt = cls_at[idx][ev_site] + (ev_cls - 1)
        np.add.at(cov, tgt[valid], ev_cnt[valid].astype(np.uint64))
        # a window with a non-ACGT character is "undefined" on the device; none occurs in aligned ACGT rows without gaps
        und = np.zeros(len(idx), dtype=bool)
        np.logical_or.at(und, e_site[bad_win], True)
        st[und] = SITE_UNDEFINED
        status[idx] = st
        checked[idx] = True
        for s in np.flatnonzero(und):   # no coverage is accumulated for an undefined site
            a0 = cls_at[idx[s]]
            cov[a0:a0 + nr[idx[s]]] = 0

    general = np.flatnonzero((~skipped) & (~simple))[:max_general]
    rows_cache = {}
    for v in general:
        b = int(b_of[v])
        n_r, L = int(nr_b[b]), int(L_b[b])
        if b not in rows_cache:
            r0 = int(rows_off[b])
            rows_cache[b] = [bytes(msa["rows"][r0 + i * L:r0 + (i + 1) * L]).decode() for i in range(n_r)]
        rows = rows_cache[b]
        part = [int(x) for x in cls[cls_at[v]:cls_at[v] + n_r]]
        checked[v] = True
        if n_r > 16:
            status[v] = SITE_UNDEFINED
            continue
        try:
            kms = caller.site_kmers(rows, int(col[v]), k, bool(kind[v] == 1), int(n_ind_before[v]))
            if any(len(s) != k or set(s) - set("ACGT") for s in kms):
                raise IndexError
        except (AssertionError, IndexError):
            status[v] = SITE_UNDEFINED
            continue
        flat = np.frombuffer("".join(kms).encode(), dtype=np.uint8)
        cnt, fnd = lookup(flat, np.arange(n_r + 1, dtype=np.uint64) * np.uint64(k))
        st, tc = caller.site_outcome(part, kms, 0, cnt, fnd, low, up)
        status[v] = st
        cov[cls_at[v]:cls_at[v] + len(tc)] = np.asarray(tc, dtype=np.uint64)
    return checked, status, ncls, cov


def compare_site_cov(sites: dict, msa: dict, checked, status, ncls, cov) -> int:
    """number of checked variable columns whose status / class count / class coverages differ"""
    n_sites = len(status)
    if len(sites["status"]) != n_sites:
        return max(n_sites, len(sites["status"]))
    bad = checked & ((sites["status"] != status) | (sites["n_class"] != ncls))
    nb = int(msa["n_bubbles"])
    nv = np.diff(msa["var_off"].astype(np.int64))
    nr = msa["n_rows"].astype(np.int64)[np.repeat(np.arange(nb), nv)]
    site_of_entry = np.repeat(np.arange(n_sites), nr)
    neq = sites["cov"][:len(cov)] != cov
    bad_cov = np.zeros(n_sites, dtype=bool)
    np.logical_or.at(bad_cov, site_of_entry[neq], True)
    bad |= checked & bad_cov
    return int(bad.sum())
