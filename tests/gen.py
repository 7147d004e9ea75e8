"""Seeded random inputs for the parity tests (bubbles for SeqAlign, sequences / databases for KMC)."""
from __future__ import annotations

import os

import numpy as np

from ploidyfrost_b200.synth import kmcdb


def rand_seq(rng, n, alphabet="ACGT") -> str:
    a = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    return bytes(a[rng.integers(0, len(a), n)]).decode()


def mutate(rng, s: str, n_snp: int, n_indel: int, alphabet="ACGT", max_indel=6) -> str:
    s = list(s)
    for _ in range(n_snp):
        if not s:
            break
        i = int(rng.integers(0, len(s)))
        c = alphabet[int(rng.integers(0, len(alphabet)))]
        s[i] = c
    for _ in range(n_indel):
        if len(s) < 4:
            break
        i = int(rng.integers(0, len(s)))
        ln = int(rng.integers(1, max_indel + 1))
        if rng.random() < 0.5:
            del s[i:i + ln]
        else:
            s[i:i] = list(rand_seq(rng, ln, alphabet))
    return "".join(s)


def sort_branching(strs):
    """Order produced by a correct sort under CDBG::sortSeq_branching's key (length desc, then bytes desc)."""
    return sorted(strs, key=lambda s: (-len(s), [-ord(c) for c in s]))


def random_bubble(rng, n_rows=None, len_range=(30, 90), alphabet="ACGT", max_snp=3, max_indel=2,
                  max_indel_len=6, distinct=True):
    n = int(n_rows if n_rows is not None else rng.choice([2, 2, 2, 3, 3, 4, 5, 6]))
    L = int(rng.integers(len_range[0], len_range[1] + 1))
    base = rand_seq(rng, L, alphabet)
    out = [base]
    tries = 0
    while len(out) < n and tries < 100:
        tries += 1
        src = out[int(rng.integers(0, len(out)))] if rng.random() < 0.3 else base
        m = mutate(rng, src, int(rng.integers(0, max_snp + 1)), int(rng.integers(0, max_indel + 1)), alphabet,
                   max_indel_len)
        if len(m) < 2:
            continue
        if distinct and m in out:
            continue
        out.append(m)
    while len(out) < 2:
        out.append(rand_seq(rng, L, alphabet))
    return sort_branching(out)


def random_bubbles(seed, n, **kw):
    rng = np.random.default_rng(seed)
    return [random_bubble(rng, **kw) for _ in range(n)]


def make_genome_db(tmpdir, seed=7, genome_len=20000, k=25, version=0x200, p=5, counter_size=2, n_bins=64,
                   sig_len=9, extra_copies=2, name=None, both_strands=True, min_count=1, max_count=10000):
    """Random genome -> canonical k-mer counts -> KMC files.  Returns (prefix, genome string, kmers, counts)."""
    rng = np.random.default_rng(seed)
    g = rand_seq(rng, genome_len)
    seqs = [g]
    for _ in range(extra_copies):  # uneven depth so counts vary
        a = int(rng.integers(0, genome_len // 2))
        b = int(rng.integers(a + k, genome_len))
        seqs.append(g[a:b])
    u, c = kmcdb.count_canonical_kmers(seqs, k)
    prefix = os.path.join(str(tmpdir), name or f"db_v{version}_p{p}_c{counter_size}")
    kmcdb.write_kmc_db(prefix, u, c, k, version=version, lut_prefix_len=p, counter_size=counter_size,
                       n_bins=n_bins, sig_len=sig_len, both_strands=both_strands, min_count=min_count,
                       max_count=max_count)
    return prefix, g, u, c


def query_sequences(rng, genome: str, n, k=25, len_range=(25, 200), p_mut=0.3, p_n=0.1, p_rc=0.5, p_short=0.05):
    """Substrings of the genome, some mutated / reverse-complemented / containing N / shorter than k."""
    comp = str.maketrans("ACGTacgt", "TGCAtgca")
    out = []
    for _ in range(n):
        if rng.random() < p_short:
            out.append(rand_seq(rng, int(rng.integers(0, k))))
            continue
        ln = int(rng.integers(len_range[0], len_range[1] + 1))
        a = int(rng.integers(0, len(genome) - ln))
        s = genome[a:a + ln]
        if rng.random() < p_mut:
            s = mutate(rng, s, int(rng.integers(1, 3)), 0)
        if rng.random() < p_n:
            i = int(rng.integers(0, len(s)))
            s = s[:i] + "N" + s[i + 1:]
        if rng.random() < p_rc:
            s = s.translate(comp)[::-1]
        if rng.random() < 0.1:
            s = s.lower()
        out.append(s)
    return out
