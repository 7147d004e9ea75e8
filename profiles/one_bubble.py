"""Diagnostics: time / profile ONE pathological bubble (a 32 k-step co-optimal DFS, region 1 of the bench genome).
   PF_GROUP_LANES=32,32,32,32,32 PF_LANE_STEP_LIMIT=100000000 PF_HEAVY_CTAS=0 python profiles/one_bubble.py [copies]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.bindings import flatten_bubbles
from ploidyfrost_b200 import capi

B = ["CTAACGATAACACCGGACGTGATTACGTATTACCTGAAGCTACCGGGGCGCCTGTTGCCAAGCGTTACGTCGATCAAGCTAGCCTTACGACCGTCTTATCGTTAACCACCGCAGCTTCGCTAGTGTGTGTATCTATGTTTTATCCTCGCCCGCCCGAGCTATCTCCACAAGACACA",
     "CTAACGATAACACCGGACGTGATTACGTATTACCTGAAGCTACCGGGGCGCCTGTTGCCAAGCGTTCCGGCGATCCAGCTAGCCTTACGACCGTCTTATCGTTAACCACCGCAGCGAGTGTGTATCTATGGTTTATCCTCGCCGGCCCGAGCTATCTCCACAAGACACA",
     "CTAACGATAACACCGGACGTGATTACGTATTACCTGAAGCTACCGGGGCGCCTGTTGCCAAGCGTTCCGGCGATCAAGCTAGCCTTACGACCGTCTTATCATTAACCACCGCAGCGAGTGTGTATCTATGTTTTATCCTCGCCCGCCCGAGCTATCTCCACAAGACACA",
     "CTAACGATAACACCGGACGTGATTAAGTATTACGTGAAGCTACCGTGGCGCCTGTTGCCAAGCGTTCCGGCGATCAAGCTAGCCTTACGACCGTCTTATCGTTAACCACCGCAGCGAGTGTGTATCTATGTTATATCCTCGCCCGCCCGAGCTATCTCCACAAGACACA"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = capi.Context(0)
flat = flatten_bubbles([B] * n)
for it in range(3):
    t = time.perf_counter()
    m = ctx.align(*flat)
    dt = time.perf_counter() - t
    print(f"{n} copies: {dt * 1e3:.2f} ms, status {set(m['status'].tolist())}, rows {m['n_rows'][0]} x {m['aln_len'][0]}, tiers {ctx.last_tier_counts}, cells {ctx.last_cells}")
ctx.close()
