"""-m gpu: the entry points added for the e2e path and the coloured caller -- pf_align_staged (align the branches the lookup
call staged on the device), pf_site_kmers (site k-mers without lookups), pf_lookup_partition (SM partition stream)."""
import numpy as np
import pytest

from oracle import caller
from oracle.bindings import flatten_bubbles, flatten_seqs
from tests import gen
from tests.util import assert_msa_equal

pytestmark = pytest.mark.gpu


def _bubble_batch(seed, n, **kw):
    bubbles = gen.random_bubbles(seed, n, **kw)
    return bubbles, flatten_bubbles(bubbles)


def test_align_staged_equals_align(gpu_ctx, oracle, tmp_path):
    from ploidyfrost_b200 import capi
    k = 25
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=9, k=k, version=0x200, p=9, genome_len=40000)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        for seed, kw in ((1, {}), (2, dict(len_range=(100, 300), max_indel_len=40)), (3, dict(alphabet="AC"))):
            bubbles, (bases, off, boff) = _bubble_batch(seed, 1500, **kw)
            rng = np.random.default_rng(seed)
            ents = gen.query_sequences(rng, g, len(bubbles), k=k)
            eb, eo = flatten_seqs(ents)
            lb = np.concatenate([eb, bases])
            lo = np.concatenate([eo, off[1:] + eo[-1]]).astype(np.uint64)
            want = gpu_ctx.align(bases, off, boff)
            cov = db.cov(lb, lo, mode=capi.LOOKUP_FWD_THEN_RC)          # stages entrances + branches on the device
            ln = np.diff(off)
            got = gpu_ctx.align_staged(db, len(ents), boff, int(ln.max()), int(np.diff(boff).max()))
            assert_msa_equal(want, got, bubbles, f"staged seed {seed}")
            assert len(cov) == len(lo) - 1
            assert_msa_equal(oracle.align(bases, off, boff, n_threads=8), got, bubbles, f"staged vs oracle seed {seed}")
        with pytest.raises(capi.PfError):                                # more sequences than were staged
            gpu_ctx.align_staged(db, len(lo), boff, 10, 2)
    finally:
        db.close()


def test_site_kmers_match_the_restatement(gpu_ctx, oracle):
    k = 25
    bubbles, flat = _bubble_batch(11, 800, len_range=(60, 140), max_indel=3, max_snp=4)
    m = gpu_ctx.align(*flat)
    sk = gpu_ctx.site_kmers(k)
    assert np.array_equal(sk["site_off"], m["var_off"]) and np.array_equal(sk["key_off"], m["cls_off"])
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    n_ok = n_und = 0
    for b in range(len(bubbles)):
        nr, L = int(m["n_rows"][b]), int(m["aln_len"][b])
        r0 = int(m["rows_off"][b])
        rows = [bytes(m["rows"][r0 + i * L:r0 + (i + 1) * L]).decode() for i in range(nr)]
        v0, v1 = int(m["var_off"][b]), int(m["var_off"][b + 1])
        n_ind = 0
        for v in range(v0, v1):
            is_ind = m["var_kind"][v] == 1
            keys = sk["keys"][int(m["cls_off"][b]) + (v - v0) * nr: int(m["cls_off"][b]) + (v - v0 + 1) * nr]
            try:
                want = caller.site_kmers(rows, int(m["var_col"][v]), k, bool(is_ind), n_ind)
                if any(len(s) != k for s in want):
                    raise IndexError
            except (AssertionError, IndexError):
                assert sk["status"][v] == 3
                n_und += 1
            else:
                assert sk["status"][v] == 0
                for s, key in zip(want, keys):
                    val = 0
                    for ch in s:
                        val = val * 4 + code[ch]
                    assert val == int(key), (b, v, s)
                n_ok += 1
            if is_ind:
                n_ind += 1
    assert n_ok > 500


def test_lookup_partition_stream(gpu_ctx, oracle, tmp_path):
    """a green-context stream: lookups enqueued on it give the same answers"""
    import torch
    from ploidyfrost_b200 import capi
    k = 25
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=5, k=k, version=0x200, p=9, genome_len=40000)
    ctx = capi.Context(0)
    db = capi.KmcDb(ctx, prefix)
    try:
        try:
            ptr, sms = ctx.lookup_partition(32)
        except capi.PfError as e:
            if "green contexts" in str(e):
                pytest.skip(str(e))
            raise
        assert ptr and 8 <= sms <= 140
        assert ctx.lookup_partition(32)[0] == ptr            # the same partition is handed out again
        rng = np.random.default_rng(1)
        bases, off = flatten_seqs(gen.query_sequences(rng, g, 3000, k=k))
        wo = capi.window_offsets(off, k)
        dev = torch.device("cuda", 0)
        pad = (-len(bases)) % 16
        d_b = torch.from_numpy(np.concatenate([bases, np.zeros(pad, np.uint8)])).to(dev)
        d_o = torch.from_numpy(off.astype(np.int64)).to(dev)
        d_w = torch.from_numpy(wo.astype(np.int64)).to(dev)
        nw = int(wo[-1])
        d_c = torch.zeros(nw, dtype=torch.int32, device=dev)
        d_f = torch.zeros(nw, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        db.lookup_dev(d_b.data_ptr(), len(bases), d_o.data_ptr(), d_w.data_ptr(), len(off) - 1, nw, capi.LOOKUP_CANONICAL, 0, 0xFFFFFFFF,
                      d_c.data_ptr(), d_f.data_ptr(), None, ptr)
        torch.cuda.synchronize()
        ho = oracle.kmc_open(prefix)
        co, fo = oracle.kmc_counts(ho, bases, off, k, mode=0, n_threads=4)
        oracle.kmc_close(ho)
        assert np.array_equal(d_c.cpu().numpy().view(np.uint32), co) and np.array_equal(d_f.cpu().numpy(), fo)
    finally:
        db.close()
        ctx.close()
