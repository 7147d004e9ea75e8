"""Pins the oracle (oracle/pf_oracle.cpp) against the unmodified reference compiled into oracle/_ref.

The reference ships no tests for this path, so these differential runs -- plus the
golden vectors generated from the same build (tests/golden) -- are what pins parity.
"""
import numpy as np
import pytest

from oracle.bindings import flatten_seqs
from tests import gen
from tests.util import assert_msa_equal

ALIGN_CASES = [
    (1, {}, {}),
    (2, dict(alphabet="AC"), {}),
    (3, dict(alphabet="AC", len_range=(10, 40), max_indel=3), {}),
    (4, dict(max_indel=4, max_snp=5), {}),
    (5, {}, dict(M=2.5, D=-1.5, G=-3.5)),
    (6, {}, dict(M=1.7, D=-0.3, G=-2.2)),
    (7, dict(alphabet="AC"), dict(M=3, D=-2, G=-1.5)),
    (8, dict(len_range=(100, 300), max_indel_len=40), {}),
    (9, dict(alphabet="A", len_range=(5, 30)), {}),
    (10, dict(alphabet="ACG", max_indel=5, max_indel_len=3), {}),
    (12, {}, dict(M=200, D=-100, G=-300)),
    (13, dict(alphabet="AC", len_range=(20, 60)), dict(M=1, D=-1, G=-1)),
]


@pytest.mark.parametrize("seed,kw,sc", ALIGN_CASES)
def test_seqalign_oracle_equals_reference(oracle, ref, seed, kw, sc):
    bubbles = gen.random_bubbles(seed, 1500, **kw)
    a = ref.align_bubbles(bubbles, n_threads=8, **sc)
    b = oracle.align_bubbles(bubbles, n_threads=8, **sc)
    assert_msa_equal(a, b, bubbles, f"seed {seed}")
    assert (a["n_rows"] == 0).sum() > 0 or seed == 9   # the empty-alignment path is exercised


KMC_CASES = [(0, 5, 2, 25), (0x200, 5, 2, 25), (0x200, 9, 2, 25), (0, 9, 1, 25), (0x200, 9, 3, 25), (0, 1, 4, 25),
             (0x200, 3, 2, 31), (0, 4, 2, 32), (0x200, 5, 2, 21), (0x200, 2, 2, 18)]


@pytest.mark.parametrize("ver,p,C,k", KMC_CASES)
def test_kmc_oracle_equals_reference(oracle, ref, tmp_path, ver, p, C, k):
    rng = np.random.default_rng(100 + p + k)
    sig = 9 if k >= 25 else 7
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=ver + p, k=k, version=ver, p=p, counter_size=C, sig_len=sig,
                                         extra_copies=3)
    hr, ho = ref.kmc_open(prefix), oracle.kmc_open(prefix)
    try:
        ir, io = ref.kmc_info(hr), oracle.kmc_info(ho)
        for f in ("kmer_length", "mode", "counter_size", "lut_prefix_length", "min_count", "max_count", "total_kmers",
                  "both_strands"):
            assert ir[f] == io[f], f
        bases, off = flatten_seqs(gen.query_sequences(rng, g, 800, k=k))
        for mode, api in ((0, True), (0, False), (1, False), (2, False)):
            cr, fr = ref.kmc_counts(hr, bases, off, k, mode=mode, use_read_api=api, n_threads=4)
            co, fo = oracle.kmc_counts(ho, bases, off, k, mode=mode, n_threads=4)
            assert np.array_equal(cr, co) and np.array_equal(fr, fo), (mode, api)
            assert fr.sum() > 0
        for mode in (0, 1, 2):
            a = ref.kmc_cov(hr, bases, off, mode=mode, low=1, up=3, n_threads=4)
            b = oracle.kmc_cov(ho, bases, off, mode=mode, low=1, up=3, n_threads=4)
            assert np.array_equal(a, b)
        for h, chk in ((hr, ref), (ho, oracle)):
            chk.kmc_set_min_count(h, 2)
            chk.kmc_set_max_count(h, 3)
        cr, fr = ref.kmc_counts(hr, bases, off, k, mode=1, use_read_api=False)
        co, fo = oracle.kmc_counts(ho, bases, off, k, mode=1)
        assert np.array_equal(cr, co) and np.array_equal(fr, fo)
    finally:
        ref.kmc_close(hr)
        oracle.kmc_close(ho)
