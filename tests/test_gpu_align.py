"""-m gpu: the CUDA SeqAlign path (through the C ABI) against the oracle, bit-exact on every output field."""
import numpy as np
import pytest

from oracle.bindings import flatten_bubbles, msa_bubble
from tests import gen
from tests.util import assert_msa_equal

pytestmark = pytest.mark.gpu

LITERALS = [  # SURVEY.md section 4, outputs of the unmodified SeqAlign
    (["ACGTACGTAC", "ACGTTCGTAC"], ["ACGTACGTAC", "ACGTTCGTAC"], [4], [], [], {4: [1, 2]}),
    (["ACGTACGGGTAC", "ACGTACGTAC"], ["ACGTACGGGTAC", "ACGTACG--TAC"], [], [7], [2], {7: [1, 2]}),
    (["ACGTACGTAC", "ACGTACGGGTAC"], ["ACGTACG--TAC", "ACGTACGGGTAC"], [], [7], [2], {7: [1, 2]}),
    (["AAAAAAAAAA", "AAAAAAAA"], ["AAAAAAAAAA", "AAAAAAAA--"], [], [8], [], {8: [1, 2]}),
    (["ACGTAAAATTGCA", "ACGTAAATTGCA", "ACGTCAAATTGCA"], ["ACGTAAAATTGCA", "ACGTAAA-TTGCA", "ACGTCAAATTGCA"], [4], [7],
     [1], {4: [1, 1, 2], 7: [1, 2, 1]}),
    (["GATTACAGATTACA", "GATTACATTACA", "GATTACAGATTCCA", "GATTACATTCCA"],
     ["GATTACAGATTACA", "GATTACA--TTACA", "GATTACAGATTCCA", "GATTACA--TTCCA"], [11], [7], [2],
     {7: [1, 2, 1, 2], 11: [1, 1, 2, 2]}),
]


def test_literal_known_answers(gpu_ctx):
    bubbles = [x[0] for x in LITERALS]
    m = gpu_ctx.align(*flatten_bubbles(bubbles))
    for i, (_, rows, snp, ind, ilen, part) in enumerate(LITERALS):
        r = msa_bubble(m, i)
        assert r["status"] == 0
        assert r["rows"] == rows and r["snp_pos"] == snp and r["indel_pos"] == ind and r["indel_len"] == ilen
        assert r["partition"] == part


CASES = [
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
    (11, dict(len_range=(49, 49), max_indel=0, max_snp=1, n_rows=2), {}),
    (12, {}, dict(M=200, D=-100, G=-300)),          # integral but too large for the s16x2 fill -> INT32 variant
    (13, dict(alphabet="AC", len_range=(20, 60)), dict(M=1, D=-1, G=-1)),
    (14, dict(len_range=(150, 250), max_indel_len=20, max_indel=3), {}),   # the 192/256 size classes
]


@pytest.mark.parametrize("seed,kw,sc", CASES)
def test_align_matches_oracle(gpu_ctx, oracle, seed, kw, sc):
    bubbles = gen.random_bubbles(seed, 3000, **kw)
    flat = flatten_bubbles(bubbles)
    a = oracle.align(*flat, n_threads=8, **sc)
    b = gpu_ctx.align(*flat, **sc)
    assert_msa_equal(a, b, bubbles, f"seed {seed}")
    assert (b["status"] == 0).all()


@pytest.mark.parametrize("lanes", ["2,4,8,16,32", "32,16,8,4,2", "1,1,1,1,1", "4,4,4,4,4"])
def test_lanes_per_bubble_configurations(oracle, lanes, monkeypatch):
    """The group kernel (G lanes per bubble, strip-pipelined fill) at every G, on every size class, against the oracle.
    PF_GROUP_LANES is read once per context, so each configuration gets its own."""
    from ploidyfrost_b200 import capi
    monkeypatch.setenv("PF_GROUP_LANES", lanes)
    ctx = capi.Context(0)
    try:
        for seed, kw in ((1, {}), (2, dict(alphabet="AC")), (8, dict(len_range=(100, 300), max_indel_len=40)),
                         (14, dict(len_range=(150, 250), max_indel_len=20, max_indel=3)), (9, dict(alphabet="A", len_range=(5, 30))),
                         (21, dict(len_range=(2, 12)))):
            bubbles = gen.random_bubbles(seed, 1200, **kw)
            flat = flatten_bubbles(bubbles)
            a = oracle.align(*flat, n_threads=8)
            b = ctx.align(*flat)
            assert_msa_equal(a, b, bubbles, f"lanes {lanes} seed {seed}")
    finally:
        ctx.close()


def test_align_matches_reference_when_available(gpu_ctx, ref):
    bubbles = gen.random_bubbles(99, 4000)
    flat = flatten_bubbles(bubbles)
    a = ref.align(*flat, n_threads=8)
    b = gpu_ctx.align(*flat)
    assert_msa_equal(a, b, bubbles, "vs reference")


def test_empty_batch_and_bad_bubbles(gpu_ctx):
    m = gpu_ctx.align(*flatten_bubbles([]))
    assert m["n_bubbles"] == 0
    m = gpu_ctx.align(*flatten_bubbles([["ACGT"], ["ACGT", "ACGA"]]))
    assert m["status"][0] == 5 and m["n_rows"][0] == 0      # PF_BUBBLE_BAD_INPUT
    assert m["status"][1] == 0 and m["n_rows"][1] == 2


def test_literal_dash_in_input_is_reported_not_guessed(gpu_ctx, oracle):
    """'-' is SeqAlign's gap character; an input that already contains one is outside the contract (no caller produces
    it) and must come back as PF_BUBBLE_BAD_INPUT, the other bubbles of the batch untouched."""
    bubbles = gen.random_bubbles(22, 600, alphabet="AC-", len_range=(20, 140)) + gen.random_bubbles(23, 600)
    flat = flatten_bubbles(bubbles)
    a = oracle.align(*flat, n_threads=8)
    b = gpu_ctx.align(*flat)
    for i, bub in enumerate(bubbles):
        if any("-" in s for s in bub):
            assert b["status"][i] == 5 and b["n_rows"][i] == 0
        else:
            assert msa_bubble(a, i) == msa_bubble(b, i)


def test_long_pair_matches_oracle(gpu_ctx, oracle):
    rng = np.random.default_rng(3)
    bubbles = []
    for L in (600, 1200, 2500):
        a = gen.rand_seq(rng, L)
        i = int(rng.integers(100, L - 200))
        b = a[:i] + a[i + 90:]
        b = gen.mutate(rng, b, 3, 0)
        bubbles.append(gen.sort_branching([a, b]))
    flat = flatten_bubbles(bubbles)
    a = oracle.align(*flat, n_threads=4)
    b = gpu_ctx.align(*flat)
    assert_msa_equal(a, b, bubbles, "long pairs")


def test_indel_heavy_5kbp_branches(gpu_ctx, oracle):
    """BASELINE configs[4]: branches up to 5 kbp with long indels, three and four rows -- beyond what the int16 fill can score,
    so the INT32 wavefront of the warp kernel and the large-limits tier carry it."""
    rng = np.random.default_rng(17)
    bubbles = []
    for L, rows in ((5000, 2), (4200, 3), (3000, 4), (1900, 3)):
        a = gen.rand_seq(rng, L)
        out = [a]
        for r in range(rows - 1):
            i = int(rng.integers(200, L - 1500))
            d = int(rng.integers(100, 1200))
            b = a[:i] + a[i + d:] if r % 2 == 0 else a[:i] + gen.rand_seq(rng, d // 4) + a[i:]
            out.append(gen.mutate(rng, b, 4, 1))
        bubbles.append(gen.sort_branching(out))
    flat = flatten_bubbles(bubbles)
    a = oracle.align(*flat, n_threads=4)
    b = gpu_ctx.align(*flat)
    assert_msa_equal(a, b, bubbles, "5 kbp branches")
    assert (b["status"] == 0).all()


def test_exploding_traceback_goes_through_the_heavy_queue(oracle):
    """A batch in which many bubbles exceed the first-pass DFS budget (two-letter alphabet: co-optimal paths abound):
    the heavy queue takes what it can beside the first pass, the pass after it takes the rest; results are the oracle's."""
    from ploidyfrost_b200 import capi
    ctx = capi.Context(0)
    try:
        bubbles = gen.random_bubbles(31, 4000, alphabet="AC", len_range=(60, 200), max_indel=4, max_indel_len=12)
        flat = flatten_bubbles(bubbles)
        a = oracle.align(*flat, n_threads=8)
        b = ctx.align(*flat)
        assert_msa_equal(a, b, bubbles, "heavy queue")
        assert ctx.last_heavy_queued > 0
    finally:
        ctx.close()


def test_idempotence_and_permutation_properties(gpu_ctx):
    """Size-independent properties: removing gaps from the aligned rows gives back the inputs, in order; all
    rows of a bubble have the same length; the batch result does not depend on batch composition."""
    bubbles = gen.random_bubbles(123, 20000, len_range=(40, 70))
    flat = flatten_bubbles(bubbles)
    m = gpu_ctx.align(*flat)
    assert (m["status"] == 0).all()
    for i in range(0, len(bubbles), 37):
        r = msa_bubble(m, i)
        if r["rows"]:
            assert [x.replace("-", "") for x in r["rows"]] == bubbles[i]
            assert len({len(x) for x in r["rows"]}) == 1
    sub = bubbles[5000:5100]
    m2 = gpu_ctx.align(*flatten_bubbles(sub))
    for i in range(100):
        assert msa_bubble(m2, i) == msa_bubble(m, 5000 + i)


def test_align_matches_golden_reference_vectors(gpu_ctx):
    """Committed outputs of the unmodified reference (tests/golden, no /root/reference needed on this box)."""
    from tests.test_cpu_golden import check_align
    check_align(lambda b, M, D, G: gpu_ctx.align(*flatten_bubbles(b), M=M, D=D, G=G))
