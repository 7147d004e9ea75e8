"""-m gpu: the CUDA KMC index + lookup kernel (through the C ABI) against the oracle, bit-exact."""
import os

import numpy as np
import pytest

from oracle.bindings import flatten_seqs
from tests import gen

pytestmark = pytest.mark.gpu

KMC_CASES = [(0, 5, 2, 25), (0x200, 5, 2, 25), (0x200, 9, 2, 25), (0, 9, 1, 25), (0x200, 9, 3, 25), (0, 1, 4, 25),
             (0x200, 3, 2, 31), (0, 4, 2, 32), (0x200, 5, 2, 21), (0x200, 2, 2, 18), (0, 5, 4, 29)]


def expect_hash(k, C):
    """pf_kmc_hash.cuh: the slot (3 distance bits + remainder + counter) must fit 63 bits with a table <= 256 MB."""
    return 2 * k + 8 * C + 3 - 63 <= 23


@pytest.mark.parametrize("index", ["auto", "verbatim"])
@pytest.mark.parametrize("ver,p,C,k", KMC_CASES)
def test_lookup_matches_oracle(gpu_ctx, oracle, tmp_path, ver, p, C, k, index):
    from ploidyfrost_b200 import capi
    rng = np.random.default_rng(200 + p + k)
    sig = 9 if k >= 25 else 7
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=ver + p + 1, k=k, version=ver, p=p, counter_size=C, sig_len=sig,
                                         extra_copies=3)
    ho = oracle.kmc_open(prefix)
    db = capi.KmcDb(gpu_ctx, prefix, index=index)
    try:
        assert db.index_kind == ("hash" if index == "auto" and expect_hash(k, C) else "verbatim")
        io = oracle.kmc_info(ho)
        for f in ("kmer_length", "mode", "counter_size", "lut_prefix_length", "min_count", "max_count", "total_kmers",
                  "both_strands", "kmc_version", "n_bins"):
            assert db.info[f] == io[f], f
        seqs = gen.query_sequences(rng, g, 3000, k=k) + ["", "A", g[:k], g[5:5 + k - 1], "N" * 40, g[:3000]]
        bases, off = flatten_seqs(seqs)
        for mode in (0, 1, 2):
            co, fo = oracle.kmc_counts(ho, bases, off, k, mode=mode, n_threads=4)
            cg, fg = db.counts(bases, off, mode=mode)
            assert np.array_equal(co, cg), mode
            assert np.array_equal(fo, fg), mode
            assert fo.sum() > 0
            a = oracle.kmc_cov(ho, bases, off, mode=mode, low=1, up=3, n_threads=4)
            b = db.cov(bases, off, mode=mode, low=1, up=3)
            assert np.array_equal(a, b), mode
        oracle.kmc_set_min_count(ho, 2); oracle.kmc_set_max_count(ho, 3)
        db.set_min_count(2); db.set_max_count(3)
        co, fo = oracle.kmc_counts(ho, bases, off, k, mode=0)
        cg, fg = db.counts(bases, off, mode=0)
        assert np.array_equal(co, cg) and np.array_equal(fo, fg)
        db.reset_min_max()
        assert db.refresh_info()["min_count"] == io["min_count"]
    finally:
        oracle.kmc_close(ho)
        db.close()


def test_lookup_against_reference_when_available(gpu_ctx, ref, tmp_path):
    from ploidyfrost_b200 import capi
    rng = np.random.default_rng(5)
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=77, k=25, version=0x200, p=9, genome_len=50000)
    hr = ref.kmc_open(prefix)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        bases, off = flatten_seqs(gen.query_sequences(rng, g, 4000, k=25))
        cr, fr = ref.kmc_counts(hr, bases, off, 25, mode=0, use_read_api=True, n_threads=4)   # GetCountersForRead
        cg, fg = db.counts(bases, off, mode=0)
        assert np.array_equal(cr, cg)
        cr, fr = ref.kmc_counts(hr, bases, off, 25, mode=1, use_read_api=False, n_threads=4)  # readCov pattern
        cg, fg = db.counts(bases, off, mode=1)
        assert np.array_equal(cr, cg) and np.array_equal(fr, fg)
    finally:
        ref.kmc_close(hr)
        db.close()


def test_empty_and_degenerate_batches(gpu_ctx, tmp_path):
    from ploidyfrost_b200 import capi
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=3, k=25, version=0, p=5, genome_len=3000)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        bases, off = flatten_seqs([])
        cg, fg = db.counts(bases, off)
        assert len(cg) == 0
        bases, off = flatten_seqs(["", "ACG", ""])
        cg, fg = db.counts(bases, off)
        assert len(cg) == 0
        cov = db.cov(bases, off)
        assert (cov["n_kmers"] == 0).all() and (cov["min"] == 10000).all() and (cov["first_missing"] == -1).all()
    finally:
        db.close()


def test_open_errors_are_reported(gpu_ctx, tmp_path):
    from ploidyfrost_b200 import capi
    with pytest.raises(capi.PfError):
        capi.KmcDb(gpu_ctx, str(tmp_path / "does_not_exist"))
    bad = tmp_path / "bad"
    (tmp_path / "bad.kmc_pre").write_bytes(b"XXXX" + b"\0" * 64 + b"XXXX")
    (tmp_path / "bad.kmc_suf").write_bytes(b"KMCSKMCS")
    with pytest.raises(capi.PfError):
        capi.KmcDb(gpu_ctx, str(bad))


def test_large_random_property(gpu_ctx, tmp_path):
    """Size-independent properties on a bigger DB: every DB k-mer is found with its count (both strands),
    and sum of counts over the genome's windows equals the sum computed from the writer's table."""
    from ploidyfrost_b200 import capi
    from ploidyfrost_b200.synth import kmcdb
    k = 25
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=11, k=k, version=0x200, p=9, genome_len=400000, n_bins=128)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        bases, off = flatten_seqs([g])
        cg, fg = db.counts(bases, off, mode=0)
        assert fg.all()
        kv = kmcdb.canonical(kmcdb.kmers_of(kmcdb.encode_bases(g), k), k)
        idx = np.searchsorted(u, kv)
        assert np.array_equal(u[idx], kv)
        assert np.array_equal(cg.astype(np.uint64), c[idx])
        comp = str.maketrans("ACGT", "TGCA")
        bases2, off2 = flatten_seqs([g.translate(comp)[::-1]])
        cg2, _ = db.counts(bases2, off2, mode=0)
        assert np.array_equal(cg2[::-1], cg)
        cov = db.cov(bases, off, mode=1, low=0, up=100000)
        assert int(cov["sum"][0]) == int(cg.sum()) and int(cov["min"][0]) == int(cg.min())
    finally:
        db.close()


def test_lookup_matches_golden_reference_vectors(gpu_ctx):
    """Committed outputs of the unmodified CKMCFile (tests/golden), both on-disk layouts, three lookup dialects."""
    from ploidyfrost_b200 import capi
    from tests.test_cpu_golden import check_kmc
    check_kmc(lambda prefix: capi.KmcDb(gpu_ctx, prefix), lambda db: db.close(),
              lambda db, b, o, k, mode: db.counts(b, o, mode=mode),
              lambda db, b, o, mode, low, up: db.cov(b, o, mode=mode, low=low, up=up))


def _swap_two_records(prefix, k, p, C, first_bucket_of=2):
    """Swap the first two records of the first prefix bucket that holds >= `first_bucket_of` records (KMC1 layout)."""
    S = (k - p) // 4
    R = S + C
    pre = open(prefix + ".kmc_pre", "rb").read()
    lut = np.frombuffer(pre[4:4 + 8 * (4 ** p)], dtype=np.uint64)
    sizes = np.diff(np.concatenate([lut, [lut[-1]]]).astype(np.int64))
    b = int(np.argmax(sizes >= first_bucket_of))
    assert sizes[b] >= first_bucket_of
    suf = bytearray(open(prefix + ".kmc_suf", "rb").read())
    a = 4 + int(lut[b]) * R
    suf[a:a + R], suf[a + R:a + 2 * R] = suf[a + R:a + 2 * R], suf[a:a + R]
    open(prefix + ".kmc_suf", "wb").write(bytes(suf))


@pytest.mark.parametrize("chunk", [None, 1, 3])
def test_database_the_reference_cannot_search_keeps_the_verbatim_index(gpu_ctx, oracle, tmp_path, monkeypatch, chunk):
    """A bucket whose suffixes are not ascending makes BinarySearch's outcome order-dependent: the hash index must
    refuse it, and the verbatim image must still reproduce whatever the reference's search returns.  chunk: records per
    chunk of the streaming open -- with 1 the offending pair always straddles a chunk seam (the carried predecessor)."""
    from ploidyfrost_b200 import capi
    if chunk:
        monkeypatch.setenv("PF_OPEN_CHUNK_RECORDS", str(chunk))
    k, p, C = 25, 5, 2
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=21, k=k, version=0, p=p, counter_size=C, genome_len=30000)
    _swap_two_records(prefix, k, p, C, first_bucket_of=3)
    ho = oracle.kmc_open(prefix)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        assert db.index_kind == "verbatim"
        bases, off = flatten_seqs([g, g[100:5000]])
        for mode in (0, 1, 2):
            co, fo = oracle.kmc_counts(ho, bases, off, k, mode=mode, n_threads=4)
            cg, fg = db.counts(bases, off, mode=mode)
            assert np.array_equal(co, cg) and np.array_equal(fo, fg), mode
        assert not fo.all()   # the swap did hide records from the binary search
    finally:
        oracle.kmc_close(ho)
        db.close()


@pytest.mark.parametrize("ver", [0, 0x200])
def test_strand_specific_database(gpu_ctx, oracle, tmp_path, ver):
    """both_strands == false: keys are stored as written, so 'as written, else reverse complement' stays two probes
    (CDBG.cpp:38-43) and the canonical shortcut must not be taken."""
    from ploidyfrost_b200 import capi
    from ploidyfrost_b200.synth import kmcdb
    k = 25
    rng = np.random.default_rng(9)
    g = gen.rand_seq(rng, 20000)
    kv = kmcdb.kmers_of(kmcdb.encode_bases(g), k)                      # forward k-mers, NOT canonicalised
    u, c = np.unique(kv, return_counts=True)
    prefix = str(tmp_path / f"fwd_{ver}")
    kmcdb.write_kmc_db(prefix, u, c.astype(np.uint64), k, version=ver, lut_prefix_len=5, counter_size=2, n_bins=32, sig_len=9,
                       both_strands=False)
    ho = oracle.kmc_open(prefix)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        assert db.info["both_strands"] == 0 and db.index_kind == "hash"
        bases, off = flatten_seqs(gen.query_sequences(rng, g, 2000, k=k))
        for mode in (0, 1, 2):
            co, fo = oracle.kmc_counts(ho, bases, off, k, mode=mode, n_threads=4)
            cg, fg = db.counts(bases, off, mode=mode)
            assert np.array_equal(co, cg) and np.array_equal(fo, fg), mode
            a = oracle.kmc_cov(ho, bases, off, mode=mode, low=0, up=3, n_threads=4)
            b = db.cov(bases, off, mode=mode, low=0, up=3)
            assert np.array_equal(a, b), mode
    finally:
        oracle.kmc_close(ho)
        db.close()


def test_dense_buckets_and_tiny_tables(gpu_ctx, oracle, tmp_path):
    """Few k-mers (table of a handful of buckets, chains across buckets) and min/max gates on the hash path."""
    from ploidyfrost_b200 import capi
    k = 25
    for n, seed in ((40, 1), (300, 2), (5000, 3)):
        prefix, g, u, c = gen.make_genome_db(tmp_path, seed=seed, k=k, version=0x200, p=5, genome_len=n + k - 1, extra_copies=0,
                                             name=f"tiny{n}", n_bins=8)
        ho = oracle.kmc_open(prefix)
        db = capi.KmcDb(gpu_ctx, prefix)
        try:
            assert db.index_kind == "hash"
            rng = np.random.default_rng(seed)
            bases, off = flatten_seqs([g, gen.rand_seq(rng, 500)] + gen.query_sequences(rng, g, 50, k=k, len_range=(25, min(60, len(g) - 1))))
            for mode in (0, 1, 2):
                co, fo = oracle.kmc_counts(ho, bases, off, k, mode=mode)
                cg, fg = db.counts(bases, off, mode=mode)
                assert np.array_equal(co, cg) and np.array_equal(fo, fg), (n, mode)
        finally:
            oracle.kmc_close(ho)
            db.close()


def test_async_cov_beside_an_alignment(gpu_ctx, oracle, tmp_path):
    """pf_kmc_cov_async + pf_kmc_wait: same records as the synchronous call, with a pf_align of the same context in between."""
    import torch
    from oracle.bindings import flatten_bubbles
    from ploidyfrost_b200 import capi
    from tests.util import assert_msa_equal
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=41, k=25, version=0x200, p=9, genome_len=60000)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        rng = np.random.default_rng(41)
        bases, off = flatten_seqs(gen.query_sequences(rng, g, 20000, k=25))
        hb = torch.from_numpy(bases).pin_memory()
        ho = torch.from_numpy(off).pin_memory()
        out = torch.zeros(len(off) - 1, dtype=torch.uint8).new_zeros((len(off) - 1) * 24).pin_memory()
        rec = out.numpy().view(capi.COV_DTYPE)
        want = db.cov(bases, off, mode=1, low=1, up=1000).copy()
        bubbles = gen.random_bubbles(5, 3000)
        flat = flatten_bubbles(bubbles)
        for _ in range(3):
            rec[:] = 0
            db.cov_async(hb.numpy(), ho.numpy(), rec, mode=1, low=1, up=1000)
            m = gpu_ctx.align(*flat)
            db.wait()
            assert np.array_equal(rec, want)
        assert_msa_equal(oracle.align(*flat, n_threads=8), m, bubbles, "align beside async lookups")
        # offsets that do not start at zero take the staging path
        sub = off[100:]
        a = db.cov(bases, sub, mode=1, low=1, up=1000)
        assert np.array_equal(a, want[100:])
    finally:
        db.close()


def test_per_colour_lookups_config3_shape(gpu_ctx, ref, tmp_path):
    """BASELINE configs[3] at lookup level: 8 samples = 8 KMC databases open side by side on one context (CCDBG.cpp:44-84), every
    branch string read against every colour's database with that colour's own cut-offs (CCDBG::readCov, CCDBG.cpp:89-124).
    Expected: the coloured rule (oracle/caller.py) over the per-k-mer answers of the UNMODIFIED CKMCFile."""
    from oracle.caller import coloured_read_cov
    from ploidyfrost_b200 import capi
    from ploidyfrost_b200.synth import kmcdb
    k, n_col = 25, 8
    rng = np.random.default_rng(83)
    base = gen.rand_seq(rng, 30000)
    shapes = [(0x200, 9, 2), (0, 5, 1), (0x200, 5, 4), (0, 9, 2), (0x200, 1, 2), (0, 5, 4), (0x200, 9, 1), (0, 1, 2)]
    cutoff, prefixes, genomes = [], [], []
    for ci, (ver, p, C) in enumerate(shapes):
        g = gen.mutate(rng, base, 150, 20)                       # every sample has its own SNPs / indels
        depth = 3 + 2 * ci
        reads = [g] * depth + [g[a:a + 4000] for a in rng.integers(0, len(g) - 4000, 6)]
        u, c = kmcdb.count_canonical_kmers(reads, k)
        prefix = os.path.join(str(tmp_path), f"colour{ci}")
        kmcdb.write_kmc_db(prefix, u, np.minimum(c, 255 if C == 1 else 60000).astype(np.uint64), k, version=ver, lut_prefix_len=p,
                           counter_size=C, n_bins=64, sig_len=9)
        prefixes.append(prefix); genomes.append(g)
        cutoff.append((depth - 1, depth + 3))                    # -C file: one (lower, upper) per colour
    seqs = []
    for g in genomes:                                            # branch strings of every sample, some foreign to the others
        seqs += [s for s in gen.query_sequences(rng, g, 250, k=k, p_mut=0.1, p_n=0.02, p_short=0.0) if len(s) >= k]
    bases, off = flatten_seqs(seqs)
    dbs = [capi.KmcDb(gpu_ctx, p) for p in prefixes]
    n_ok = n_bad = 0
    try:
        for ci in range(n_col):
            low, up = cutoff[ci]
            hr = ref.kmc_open(prefixes[ci])
            cr, fr = ref.kmc_counts(hr, bases, off, k, mode=1, use_read_api=False, n_threads=4)
            ref.kmc_close(hr)
            cov = dbs[ci].cov(bases, off, mode=capi.LOOKUP_FWD_THEN_RC, low=low, up=up)
            w0 = 0
            for si, s in enumerate(seqs):
                n = len(s) - k + 1
                want = coloured_read_cov(cr[w0:w0 + n], fr[w0:w0 + n], low, up)
                w0 += n
                r = cov[si]
                ok = r["first_missing"] < 0 and r["first_outside"] < 0
                # the first failure, whichever kind, ends the read; on success every window was found and counted
                got = (float(r["sum"]) / float(r["n_kmers"]), True) if ok else (0.0, False)
                assert got == want, (ci, si)
                n_ok += ok; n_bad += not ok
            assert w0 == len(cr)
    finally:
        for d in dbs:
            d.close()
    assert n_ok > 100 and n_bad > 1000


@pytest.mark.parametrize("ver,p,C,k,chunk", [(0x200, 5, 2, 25, 1000), (0, 5, 2, 25, 777), (0x200, 9, 3, 25, 1), (0, 3, 4, 31, 4096)])
def test_streaming_open_chunk_seams(gpu_ctx, oracle, tmp_path, monkeypatch, ver, p, C, k, chunk):
    """pf_kmc_open streams the records through fixed-size chunks (kmc_stream_hash); with tiny chunks every seam case is hit:
    the carried predecessor of a chunk's first record (ascending-suffix check), chunks that end inside a prefix bucket, a last
    chunk that is not full.  Same answers as the oracle, whole database and partitions."""
    from ploidyfrost_b200 import capi
    monkeypatch.setenv("PF_OPEN_CHUNK_RECORDS", str(chunk))
    rng = np.random.default_rng(chunk)
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=chunk, k=k, version=ver, p=p, counter_size=C, sig_len=9 if k >= 25 else 7,
                                         genome_len=30000)
    ho = oracle.kmc_open(prefix)
    seqs = gen.query_sequences(rng, g, 2000, k=k) + [g[:2500]]
    bases, off = flatten_seqs(seqs)
    co, fo = oracle.kmc_counts(ho, bases, off, k, mode=0, n_threads=4)
    oracle.kmc_close(ho)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        if expect_hash(k, C):
            assert db.index_kind == "hash" and db.local_kmers == db.info["total_kmers"]
        cg, fg = db.counts(bases, off, mode=0)
        assert np.array_equal(co, cg) and np.array_equal(fo, fg)
    finally:
        db.close()
    if expect_hash(k, C):
        parts = [capi.KmcDb(gpu_ctx, prefix, part=r, n_parts=3) for r in range(3)]
        try:
            assert all(d.index_kind == "hash" for d in parts)
            assert sum(d.local_kmers for d in parts) == parts[0].info["total_kmers"]
        finally:
            for d in parts:
                d.close()


def test_shared_handles_on_other_contexts(gpu_ctx, oracle, tmp_path):
    """pf_kmc_share: one index in HBM, one handle + pf_ctx per host thread; four threads run different batches at the same time
    and each gets the oracle's answer."""
    import threading
    from ploidyfrost_b200 import capi
    k = 25
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=4, k=k, version=0x200, p=9, genome_len=60000)
    ho = oracle.kmc_open(prefix)
    db = capi.KmcDb(gpu_ctx, prefix)
    ctxs = [capi.Context(0) for _ in range(4)]
    shared = [capi.KmcDb(cx, prefix, share_of=db) for cx in ctxs]
    assert all(s.device_bytes == db.device_bytes and s.index_kind == db.index_kind for s in shared)
    batches, want, got = [], [], [None] * 4
    for t in range(4):
        rng = np.random.default_rng(100 + t)
        bases, off = flatten_seqs(gen.query_sequences(rng, g, 4000, k=k))
        batches.append((bases, off))
        want.append((oracle.kmc_counts(ho, bases, off, k, mode=1, n_threads=4), oracle.kmc_cov(ho, bases, off, mode=1, low=1, up=3, n_threads=4)))
    oracle.kmc_close(ho)

    def work(t):
        for _ in range(5):
            got[t] = (shared[t].counts(*batches[t], mode=1), shared[t].cov(*batches[t], mode=1, low=1, up=3))

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    try:
        for t in range(4):
            assert np.array_equal(want[t][0][0], got[t][0][0]) and np.array_equal(want[t][0][1], got[t][0][1])
            assert np.array_equal(want[t][1], got[t][1])
    finally:
        for s in shared:
            s.close()
        for cx in ctxs:
            cx.close()
        db.close()
