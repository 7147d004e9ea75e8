import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from tests import gen
from oracle.bindings import Checker, flatten_bubbles, msa_bubble
from ploidyfrost_b200 import capi
o = Checker("oracle")
cases = ((1, {}), (2, dict(alphabet="AC")), (8, dict(len_range=(100, 300), max_indel_len=40)),
         (14, dict(len_range=(150, 250), max_indel_len=20, max_indel=3)), (9, dict(alphabet="A", len_range=(5, 30))),
         (21, dict(len_range=(2, 12))), (22, dict(alphabet="AC-", len_range=(20, 140))))
for lanes in sys.argv[1:]:
    os.environ["PF_GROUP_LANES"] = lanes
    ctx = capi.Context(0)
    for seed, kw in cases:
        bubbles = gen.random_bubbles(seed, 1200, **kw)
        flat = flatten_bubbles(bubbles)
        a = o.align(*flat, n_threads=8)
        b = ctx.align(*flat)
        bad = [i for i in range(len(bubbles)) if msa_bubble(a, i) != msa_bubble(b, i)]
        print(lanes, seed, "mismatches", len(bad), "tiers", ctx.last_tier_counts)
        for i in bad[:3]:
            print("   bubble", i, [len(x) for x in bubbles[i]], "exp rows", len(msa_bubble(a, i)["rows"]), "got", msa_bubble(b, i)["status"], len(msa_bubble(b, i)["rows"]))
            if max(len(x) for x in bubbles[i]) < 30: print("   ", bubbles[i], msa_bubble(a,i)["rows"], msa_bubble(b,i)["rows"])
    ctx.close()
