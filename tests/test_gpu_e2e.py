"""-m gpu: the CUDA path (through the C ABI) against the unmodified reference's own END-TO-END output (tests/golden/e2e):
SequenceAlignment of the bubbles PloidyFrost aligned, lookup-A (readCov of every branch), lookup-B (site k-mers of the branching
bubbles) and the caller arithmetic must give back its aligned rows and every row of its coverage files byte for byte."""
import os

import numpy as np
import pytest

from oracle.bindings import flatten_seqs
from tests import e2e_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("index", ["auto", "verbatim"])
def test_reference_rows_from_the_cuda_path(gpu_ctx, index):
    from ploidyfrost_b200 import capi
    db = capi.KmcDb(gpu_ctx, os.path.join(e2e_rows.E2E, "db"), index=index)
    try:
        def kmer_counts(kmers):
            b, off = flatten_seqs(kmers)
            return db.counts(b, off, mode=capi.LOOKUP_FWD_THEN_RC)
        n_rows, n_branching = e2e_rows.check_against_reference(
            lambda *f: gpu_ctx.align(*f), lambda b, off: db.cov(b, off, mode=capi.LOOKUP_FWD_THEN_RC, low=2, up=1000), kmer_counts)
        assert n_rows == 493 and n_branching == 114
    finally:
        db.close()


@pytest.mark.parametrize("index", ["auto", "verbatim"])
def test_reference_rows_with_lookup_phase_b_on_the_device(gpu_ctx, index):
    """Same pin, but the site k-mers of the branching bubbles are built and looked up on the device (pf_site_cov) from the aligned
    rows that pf_align left in HBM -- no k-mer strings cross the host."""
    from ploidyfrost_b200 import capi
    meta, bubbles, _, _ = e2e_rows.load_fixture()
    db = capi.KmcDb(gpu_ctx, os.path.join(e2e_rows.E2E, "db"), index=index)
    try:
        site_cov, state = e2e_rows.device_site_cov_hook(db, meta, bubbles)
        n_rows, n_branching = e2e_rows.check_against_reference(
            lambda *f: gpu_ctx.align(*f), lambda b, off: db.cov(b, off, mode=capi.LOOKUP_FWD_THEN_RC, low=2, up=1000), None, site_cov)
        assert n_rows == 493 and n_branching == 114
        sc = state["sc"]
        strict_sites = [int(sc["site_off"][i]) for i, b in enumerate(bubbles) if b["strict"] and sc["site_off"][i + 1] > sc["site_off"][i]]
        assert all(sc["status"][v] == capi.SITE_SKIPPED for v in strict_sites)
    finally:
        db.close()


LIVE_CASES = [(2, 200000, 0.001, 0.01),      # BASELINE configs[0] (scaled; PF_E2E_GENOME = 1000000 for the full size)
              (4, 300000, 0.003, 0.01),      # tetraploid, three times the indel rate: 3-8 branches, sites after indels
              (2, 300000, 0.005, 0.02)]      # variant-dense diploid: bubbles of up to 8 paths


@pytest.mark.parametrize("hap,genome,p_indel,p_snp", LIVE_CASES)
def test_config0_reference_run_from_the_cuda_path(gpu_ctx, tmp_path, hap, genome, p_indel, p_snp):
    """BASELINE configs[0] (diploid, 30x reads, k = 25; scaled to 200 kbp, PF_E2E_GENOME overrides): the unmodified PloidyFrost and
    Bifrost binaries of oracle/_ref run end to end on this box, then lookup-A, SequenceAlignment and lookup-B (pf_site_cov) of the
    CUDA path must regenerate every aligned row and every coverage row the reference wrote."""
    from ploidyfrost_b200 import capi
    if e2e_rows.reference_binaries() is None:
        pytest.skip("oracle/_ref/PloidyFrost not built (make -C oracle ref_full in the dev container)")
    out, dbp = e2e_rows.run_reference_config0(str(tmp_path), genome=int(os.environ.get("PF_E2E_GENOME", str(genome))) if hap == 2 and p_indel == 0.001 else genome,
                                              haplotypes=hap, p_indel=p_indel, p_snp=p_snp, depth=15 * hap)
    meta, bubbles, _, _ = e2e_rows.load_fixture(out)
    db = capi.KmcDb(gpu_ctx, dbp)
    try:
        site_cov, state = e2e_rows.device_site_cov_hook(db, meta, bubbles)
        n_rows, n_branching = e2e_rows.check_against_reference(
            lambda *f: gpu_ctx.align(*f), lambda b, off: db.cov(b, off, mode=capi.LOOKUP_FWD_THEN_RC, low=2, up=1000), None, site_cov,
            fixture_dir=out)
        assert n_rows > 1000 and n_branching > 50, (n_rows, n_branching)
    finally:
        db.close()


def test_site_cov_statuses_against_the_host_logic(gpu_ctx, oracle, tmp_path):
    """pf_site_cov on synthetic bubbles whose variant alleles are only partly in the database and with a tight (low, up):
    every site's status (ok / dropped / missing) and class coverages must equal the reference's iteration order as restated
    in tests/e2e_rows.py (site_kmers + site_outcome) over oracle lookups."""
    from oracle.bindings import flatten_bubbles, msa_bubble
    from ploidyfrost_b200 import capi
    from ploidyfrost_b200.synth import kmcdb
    from tests import gen
    k = 25
    rng = np.random.default_rng(77)
    g = gen.rand_seq(rng, 60000)
    bubbles, extra = [], []
    for _ in range(1500):
        a = int(rng.integers(0, len(g) - 200))
        ln = int(rng.integers(2 * k + 5, 150))
        base = g[a:a + ln]
        rows = [base]
        for r in range(int(rng.integers(1, 4))):
            inner = gen.mutate(rng, base[k:-k], int(rng.integers(1, 3)), int(rng.integers(0, 2)), max_indel=4)
            alt = base[:k] + inner + base[-k:]
            if alt not in rows and len(alt) >= 2 * k:
                rows.append(alt)
                if rng.random() < 0.6:
                    extra.append(alt)                      # this allele is in the database, the others are not
        if len(rows) >= 2:
            bubbles.append(gen.sort_branching(rows))
    u, c = kmcdb.count_canonical_kmers([g] + extra, k)
    c = (c + rng.integers(0, 4, len(c))).astype(np.uint64)  # counts 1..: a (1, 4) gate drops some sites
    prefix = str(tmp_path / "sites")
    kmcdb.write_kmc_db(prefix, u, c, k, version=0x200, lut_prefix_len=5, counter_size=2, n_bins=32, sig_len=9)
    low, up = 1, 4
    h = oracle.kmc_open(prefix)
    db = capi.KmcDb(gpu_ctx, prefix)
    try:
        flat = flatten_bubbles(bubbles)
        msa = gpu_ctx.align(*flat)
        sc = db.site_cov(low, up, None)
        aligned = [msa_bubble(msa, i) for i in range(len(bubbles))]
        fake = [dict(strict=False) for _ in bubbles]
        plan, kmers = e2e_rows.branching_site_plan(fake, aligned, k)
        b, off = flatten_seqs(kmers)
        counts, found = oracle.kmc_counts(h, b, off, k, mode=1)
        seen = np.zeros(5, int)
        for bi, i, col, is_ind, n_ind, k0 in plan:
            part = aligned[bi]["partition"][col]
            st, tc = e2e_rows.site_outcome(part, kmers, k0, counts, found, low, up)
            v = int(sc["site_off"][bi]) + i
            assert int(sc["status"][v]) == st, (bi, i, int(sc["status"][v]), st)
            assert int(sc["n_class"][v]) == max(part)
            if st == 0:
                c0 = int(sc["cov_off"][bi]) + i * int(msa["n_rows"][bi])
                assert [int(x) for x in sc["cov"][c0:c0 + max(part)]] == tc
            seen[st] += 1
        assert seen[0] > 100 and seen[1] > 20 and seen[2] > 100, seen
    finally:
        oracle.kmc_close(h)
        db.close()
