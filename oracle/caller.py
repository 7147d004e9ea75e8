"""oracle/caller.py -- TEST INFRASTRUCTURE: CPU restatement of the per-site part of PloidyFrost's per-bubble caller.

Pure Python (small cases only: the end-to-end fixtures hold a few hundred bubbles).  Each function cites the reference lines it
follows; the whole is pinned against the unmodified reference's own output files by tests/e2e_rows.py (tests/golden/e2e and the
live BASELINE configs[0] runs).  Only tests/ and __graft_entry__.smoke() import this module; the product never does.
"""
from __future__ import annotations


def fmt(x: float) -> str:
    """ostream << double at default precision == printf %g (6 significant digits)."""
    return "%g" % x


def site_kmers(rows, c, k, is_indel, n_indel_before):
    """The k-mer each row contributes at variable column c (CDBG.cpp:2338-2388 indel sites, :2433-2472 SNP sites)."""
    n = len(rows)
    if is_indel:
        cur = [c] * n
        ext = [""] * n
        while True:                                       # :2338-2357: extend every row by its next base until they differ
            chars = set()
            for r in range(n):
                while rows[r][cur[r]] == "-":
                    cur[r] += 1
                ch = rows[r][cur[r]]
                cur[r] += 1
                ext[r] += ch
                chars.add(ch)
            if len(chars) > 1:
                break
        out = []
        for r in range(n):
            e = len(ext[r])
            if n_indel_before == 0:                       # :2358-2365
                start = c - k + e
                assert start >= 0, "substr with a negative start throws in the reference"
                out.append(rows[r][start:start + k - e] + ext[r])
            else:                                         # :2366-2388
                t = rows[r][:c].replace("-", "")
                if len(t) < k - e:
                    s = t + ext[r]
                    x = cur[r]
                    while len(s) < k:
                        if rows[r][x] != "-":
                            s += rows[r][x]
                        x += 1
                    out.append(s)
                else:
                    out.append(t[len(t) - (k - e):] + ext[r])
        return out
    if n_indel_before > 0:                                # :2433-2465
        out = []
        for r in range(n):
            t = rows[r][:c + 1].replace("-", "")
            if len(t) < k:
                s = t
                x = c + 1
                while len(s) < k:
                    if rows[r][x] != "-":
                        s += rows[r][x]
                    x += 1
                out.append(s)
            else:
                out.append(t[len(t) - k:])
        return out
    assert c - k + 1 >= 0
    return [rows[r][c - k + 1:c + 1] for r in range(n)]    # :2469-2472


def var_distance(i, var_site, ent_size, exit_size):
    """CDBG.cpp:2312-2330 (same in the strict path)."""
    if i == 0:
        return min(var_site[1] - var_site[0] - 1, ent_size) if len(var_site) > 1 else min(ent_size, exit_size)
    if i == len(var_site) - 1:
        return min(var_site[i] - var_site[i - 1] - 1, exit_size)
    return min(var_site[i] - var_site[i - 1] - 1, var_site[i + 1] - var_site[i] - 1)


def class_coverage(part, kmers, k0, counts, found, low, up):
    """Coverage per allele class of one site (CDBG.cpp:2393-2418): distinct k-mers per class in std::set order; returns None when
    the site is dropped."""
    sets = [dict() for _ in range(max(part))]
    for j, cl in enumerate(part):
        sets[cl - 1].setdefault(kmers[k0 + j], k0 + j)
    tc = []
    for st in sets:
        acc = 0.0
        for s in sorted(st):
            idx = st[s]
            assert found[idx], f"k-mer {s} missing: the reference exits here (CDBG.cpp:54)"
            cval = int(counts[idx])
            if not (low < cval < up):
                return None
            acc += float(cval)
        tc.append(acc)
    return tc


def site_outcome(part, kmers, k0, counts, found, low, up):
    """(status, class coverages) of one site in the reference's iteration order (classes ascending, distinct k-mers of a class in
    std::set order): 0 ok, 1 dropped at the first counter outside (low, up), 2 a missing k-mer reached first (the reference exits)."""
    sets = [dict() for _ in range(max(part))]
    for j, cl in enumerate(part):
        sets[cl - 1].setdefault(kmers[k0 + j], k0 + j)
    tc = [0] * len(sets)
    for q, st in enumerate(sets):
        for s in sorted(st):
            idx = st[s]
            if not found[idx]:
                return 2, tc
            cval = int(counts[idx])
            if not (low < cval < up):
                return 1, tc
            tc[q] += cval
    return 0, tc


def coloured_read_cov(counts, found, low, up):
    """CCDBG::readCov(s, low, up, colour) / readCovUni (CCDBG.cpp:89-124, :125-160) over the per-window answers of colour's
    database looked up 'as written, else reverse complement' (IsKmer / reverse / CheckKmer, :99-103): the windows are read in order;
    the first one that is absent, or whose count is not strictly inside (low, up), ends the read with (0, False) -- unlike the
    single-sample form (CDBG.cpp:52-56) a missing k-mer does not end the program here, the exit() after the return is dead code
    (:117).  Otherwise (sum / #windows, True).  Both-strands databases only (:95)."""
    total = 0.0
    for c, f in zip(counts, found):
        if not f:
            return 0.0, False
        if not (low < int(c) < up):
            return 0.0, False
        total += float(int(c))
    return total / len(counts), True
