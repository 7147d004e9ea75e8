"""-m gpu: large differentials on the BENCHMARKED distributions: the pfsynth workloads of bench.py run
through the C ABI and through the CPU checker (the unmodified reference when oracle/_ref is present, else our restatement),
every output field of every bubble compared -- 131 072 bubbles of the tetraploid / hexaploid shapes with all three phases
(lookup-A records, SequenceAlignment, lookup-B class coverages), and 1 200 long bubbles (50 bp .. 5 kbp branches with long
indels, 2-4 rows: BASELINE configs[4]), where the rare traceback rules get their differential cases."""
import os

import numpy as np
import pytest

from oracle import parity
from ploidyfrost_b200.synth import workload as wl

pytestmark = pytest.mark.gpu
K = 25


def _checker():
    from oracle.bindings import Checker
    try:
        return Checker("ref")
    except Exception:
        return Checker("oracle")


@pytest.mark.parametrize("n_hap,root,n_bubbles", [(4, 0, 131072), (2, 77, 65536)])
def test_pfsynth_workload_all_three_phases(gpu_ctx, tmp_path, n_hap, root, n_bubbles):
    from ploidyfrost_b200 import capi
    G = 12_000_000 if n_hap == 4 else 7_000_000
    w = wl.Workload(20261017 + n_hap, G, n_hap, p_snp=0.01, p_indel=0.001, n_threads=8, root_seed=root, divergence=0.0527 if root else 0.0)
    bb = w.bubbles(K, 6000, G - 6000, n_bubbles)
    assert bb.n_bubbles == n_bubbles
    prefix = str(tmp_path / "db")
    wl.write_db_torch(prefix, [w.haplotype(i) for i in range(n_hap)], K, 12.6, 3, device="cuda:0")
    w.close()
    ref = _checker()
    cores = os.cpu_count() or 8
    db = capi.KmcDb(gpu_ctx, prefix)
    assert db.index_kind == "hash"
    h = ref.kmc_open(prefix)
    try:
        lb, lo = bb.lookup_sequences()
        cov = db.cov(lb, lo, mode=capi.LOOKUP_FWD_THEN_RC, low=2, up=1000)
        assert parity.compare_cov(cov, ref.kmc_cov(h, lb, lo, mode=1, low=2, up=1000, n_threads=cores)) == 0
        msa = gpu_ctx.align(bb.bases, bb.seq_off, bb.bubble_off)
        msa_ref = ref.align(bb.bases, bb.seq_off, bb.bubble_off, n_threads=cores)
        bad = parity.compare_msa(msa, msa_ref)
        assert not bad.any(), f"{int(bad.sum())} bubbles differ, first {int(np.flatnonzero(bad)[0])}"
        assert (msa["status"] == 0).all()
        skip = np.ascontiguousarray(bb.bubble_type.astype(np.uint8))
        sites = db.site_cov(2, 1000, skip)
        lookup = lambda b, o: ref.kmc_counts(h, b, o, K, mode=1, use_read_api=False, n_threads=cores)
        checked, st, ncls, cv = parity.expected_site_cov(msa_ref, skip, K, 2, 1000, lookup, max_general=3000)
        assert checked.sum() > 0.5 * len(st)
        assert parity.compare_site_cov(sites, msa_ref, checked, st, ncls, cv) == 0
    finally:
        ref.kmc_close(h)
        db.close()


def test_long_bubbles_differential(gpu_ctx):
    """configs[4] shape: 1 200 bubbles, branches 50 bp .. 5 kbp (log-uniform), long indels, 2-4 rows."""
    w = wl.Workload(5, 1000, 1, n_threads=1)
    parts = [w.long_bubbles(11, K, 1000, 50, 1500, 4), w.long_bubbles(12, K, 200, 500, 5000, 4)]
    bb = wl.BubbleBatch.concat(parts)
    w.close()
    assert bb.n_bubbles == 1200 and int(np.diff(bb.seq_off).max()) > 4000
    ref = _checker()
    msa_ref = ref.align(bb.bases, bb.seq_off, bb.bubble_off, n_threads=os.cpu_count() or 8)
    msa = gpu_ctx.align(bb.bases, bb.seq_off, bb.bubble_off)
    bad = parity.compare_msa(msa, msa_ref)
    assert not bad.any(), f"{int(bad.sum())} bubbles differ, first {int(np.flatnonzero(bad)[0])}"
    assert (msa["status"] == 0).all()
