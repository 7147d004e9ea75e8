#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates the golden fixtures from the UNMODIFIED reference.

Run in the dev container (needs /root/reference, i.e. oracle/_ref/libpfref.so built by `make -C oracle ref`):

    python tests/golden/make_golden.py

The reference ships no tests or vectors for this path, so these files -- outputs of
the reference's own SeqAlign::SequenceAlignment (src/SeqAlign.cpp:550) and CKMCFile::{CheckKmer,
GetCountersForRead} (KMC/kmc_api/kmc_file.cpp:330, :904) on fixed inputs -- are what pins the oracle and the CUDA
path on machines where the reference tree is absent (the GPU box).  Everything is seeded; rerunning
reproduces the files byte for byte.

  seqalign_literals.json   the literal known-answer cases of SURVEY.md section 4 (+ the 2.5/-1.5/-3.5 scoring case)
  seqalign_random.json     600 random bubbles (4 groups: ACGT, 2-letter, fractional scoring, long indels)
  kmc_v0.kmc_pre/.kmc_suf  a 3 kbp genome's k-mers in the KMC1 layout (k=25, p=5, 2-byte counters)
  kmc_v200.kmc_pre/.kmc_suf  the same content in the KMC2 layout (signature length 7, 16 bins, p=5)
  kmc_queries.json         query sequences + the reference's answers in the three lookup dialects, both DBs
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.bindings import Checker, flatten_bubbles, flatten_seqs, msa_bubble  # noqa: E402
from tests import gen  # noqa: E402

LITERALS = [
    (["ACGTACGTAC", "ACGTTCGTAC"], (2, -1, -3)),
    (["ACGTACGGGTAC", "ACGTACGTAC"], (2, -1, -3)),
    (["ACGTACGTAC", "ACGTACGGGTAC"], (2, -1, -3)),
    (["AAAAAAAAAA", "AAAAAAAA"], (2, -1, -3)),
    (["ACGTAAAATTGCA", "ACGTAAATTGCA", "ACGTCAAATTGCA"], (2, -1, -3)),
    (["GATTACAGATTACA", "GATTACATTACA", "GATTACAGATTCCA", "GATTACATTCCA"], (2, -1, -3)),
    (["ACGTACGTAC", "ACGTTCGTAC"], (2.5, -1.5, -3.5)),
    # a real bubble of the config-2 workload whose co-optimal traceback takes ~32 000 DFS steps (100x the typical 2 x length):
    # it overflows the thread-per-bubble tier's step budget and is re-run with the flag matrix in shared memory
    (["CTAACGATAACACCGGACGTGATTACGTATTACCTGAAGCTACCGGGGCGCCTGTTGCCAAGCGTTACGTCGATCAAGCTAGCCTTACGACCGTCTTATCGTTAACCACCGCAGCTTCGCTAGTGTGTGTATCTATGTTTTATCCTCGCCCGCCCGAGCTATCTCCACAAGACACA", "CTAACGATAACACCGGACGTGATTACGTATTACCTGAAGCTACCGGGGCGCCTGTTGCCAAGCGTTCCGGCGATCCAGCTAGCCTTACGACCGTCTTATCGTTAACCACCGCAGCGAGTGTGTATCTATGGTTTATCCTCGCCGGCCCGAGCTATCTCCACAAGACACA", "CTAACGATAACACCGGACGTGATTACGTATTACCTGAAGCTACCGGGGCGCCTGTTGCCAAGCGTTCCGGCGATCAAGCTAGCCTTACGACCGTCTTATCATTAACCACCGCAGCGAGTGTGTATCTATGTTTTATCCTCGCCCGCCCGAGCTATCTCCACAAGACACA", "CTAACGATAACACCGGACGTGATTAAGTATTACGTGAAGCTACCGTGGCGCCTGTTGCCAAGCGTTCCGGCGATCAAGCTAGCCTTACGACCGTCTTATCGTTAACCACCGCAGCGAGTGTGTATCTATGTTATATCCTCGCCCGCCCGAGCTATCTCCACAAGACACA"], (2, -1, -3)),
]

RANDOM_GROUPS = [
    ("acgt", 101, dict(), (2, -1, -3)),
    ("two_letter", 102, dict(alphabet="AC", len_range=(10, 40), max_indel=3), (2, -1, -3)),
    ("fractional", 103, dict(), (1.7, -0.3, -2.2)),
    ("long_indel", 104, dict(len_range=(100, 260), max_indel_len=40), (2, -1, -3)),
]


def dump_msa(m, n):
    out = []
    for i in range(n):
        b = msa_bubble(m, i)
        b["partition"] = {str(k): v for k, v in b["partition"].items()}
        out.append(b)
    return out


def main():
    ref = Checker("ref")
    lit = []
    for seqs, (M, D, G) in LITERALS:
        m = ref.align_bubbles([seqs], M=M, D=D, G=G)
        lit.append({"input": seqs, "M": M, "D": D, "G": G, "expect": dump_msa(m, 1)[0]})
    json.dump(lit, open(os.path.join(HERE, "seqalign_literals.json"), "w"), indent=1)

    groups = []
    for name, seed, kw, (M, D, G) in RANDOM_GROUPS:
        bubbles = gen.random_bubbles(seed, 150, **kw)
        m = ref.align_bubbles(bubbles, M=M, D=D, G=G, n_threads=4)
        groups.append({"name": name, "M": M, "D": D, "G": G, "bubbles": bubbles, "expect": dump_msa(m, len(bubbles))})
    json.dump(groups, open(os.path.join(HERE, "seqalign_random.json"), "w"))

    k = 25
    res = {"k": k, "dbs": {}}
    rng = np.random.default_rng(7)
    queries = None
    for tag, ver, sig, bins in (("kmc_v0", 0, 0, 1), ("kmc_v200", 0x200, 7, 16)):
        prefix, g, u, c = gen.make_genome_db(HERE, seed=5, genome_len=3000, k=k, version=ver, p=5, counter_size=2, n_bins=bins,
                                             sig_len=sig or 9, extra_copies=3, name=tag)
        if queries is None:
            queries = gen.query_sequences(rng, g, 120, k=k, len_range=(25, 120)) + ["", "ACGT", g[:k], "N" * 30, g[100:400]]
        bases, off = flatten_seqs(queries)
        h = ref.kmc_open(prefix)
        info = ref.kmc_info(h)
        ans = {"info": {kk: int(v) for kk, v in info.items()}}
        c_read, _ = ref.kmc_counts(h, bases, off, k, mode=0, use_read_api=True)          # GetCountersForRead
        ans["GetCountersForRead"] = [int(x) for x in c_read]
        for mode, name in ((0, "canonical"), (1, "fwd_then_rc"), (2, "fwd")):
            cc, ff = ref.kmc_counts(h, bases, off, k, mode=mode, use_read_api=False)     # CheckKmer call patterns
            ans[name] = {"counts": [int(x) for x in cc], "found": [int(x) for x in ff]}
        cov = ref.kmc_cov(h, bases, off, mode=1, low=1, up=3)                          # readCov (CDBG.cpp:29-120)
        ans["readCov_low1_up3"] = [[int(r[f]) for f in ("sum", "min", "n_kmers", "first_missing", "first_outside")] for r in cov]
        ref.kmc_close(h)
        res["dbs"][tag] = ans
    res["queries"] = queries
    json.dump(res, open(os.path.join(HERE, "kmc_queries.json"), "w"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
