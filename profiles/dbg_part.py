import sys, os, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np
from tests import gen
from ploidyfrost_b200 import capi
ctx = capi.Context(0)
with tempfile.TemporaryDirectory() as d:
    for nb in (37, 64):
        prefix, g, u, c = gen.make_genome_db(d, seed=36, k=25, version=0x200, p=5, genome_len=60000, n_bins=nb, name=f"x{nb}")
        db = capi.KmcDb(ctx, prefix)
        print("n_bins", nb, db.index_kind, "build_status", db.build_status, db.info["n_bins"])
        db.close()
        db = capi.KmcDb(ctx, prefix, part=0, n_parts=1)
        print("  part 0/1:", db.index_kind, db.build_status)
        db.close()
