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


def test_config0_reference_run_from_the_cuda_path(gpu_ctx, tmp_path):
    """BASELINE configs[0] (diploid, 30x reads, k = 25; scaled to 200 kbp, PF_E2E_GENOME overrides): the unmodified PloidyFrost and
    Bifrost binaries of oracle/_ref run end to end on this box, then lookup-A, SequenceAlignment and lookup-B (pf_site_cov) of the
    CUDA path must regenerate every aligned row and every coverage row the reference wrote."""
    from ploidyfrost_b200 import capi
    if e2e_rows.reference_binaries() is None:
        pytest.skip("oracle/_ref/PloidyFrost not built (make -C oracle ref_full in the dev container)")
    out, dbp = e2e_rows.run_reference_config0(str(tmp_path), genome=int(os.environ.get("PF_E2E_GENOME", "200000")))
    meta, bubbles, _, _ = e2e_rows.load_fixture(out)
    db = capi.KmcDb(gpu_ctx, dbp)
    try:
        site_cov, state = e2e_rows.device_site_cov_hook(db, meta, bubbles)
        n_rows, n_branching = e2e_rows.check_against_reference(
            lambda *f: gpu_ctx.align(*f), lambda b, off: db.cov(b, off, mode=capi.LOOKUP_FWD_THEN_RC, low=2, up=1000), None, site_cov,
            fixture_dir=out)
        assert n_rows > 1000 and n_branching > 50
    finally:
        db.close()
