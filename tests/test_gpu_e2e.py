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
