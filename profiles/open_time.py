import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.getcwd())
t0 = time.perf_counter()
from ploidyfrost_b200 import capi
from ploidyfrost_b200.synth import kmcdb
print("import", round(time.perf_counter() - t0, 3))
tmp = tempfile.mkdtemp()
rng = np.random.default_rng(1)
for n in (2_000_000, 8_000_000):
    u = np.unique(rng.integers(0, 1 << 50, n, dtype=np.uint64))
    c = rng.integers(1, 100, len(u)).astype(np.uint64)
    p = os.path.join(tmp, f"db{n}")
    t0 = time.perf_counter(); kmcdb.write_kmc_db(p, u, c, 25, version=0x200, lut_prefix_len=5, counter_size=2, n_bins=64, sig_len=9); print("write", n, round(time.perf_counter() - t0, 3))
t0 = time.perf_counter(); ctx = capi.Context(0); print("pf_init", round(time.perf_counter() - t0, 3))
for rep in range(2):
    for n in (2_000_000, 8_000_000):
        for kind in ("auto", "verbatim"):
            t0 = time.perf_counter(); db = capi.KmcDb(ctx, os.path.join(tmp, f"db{n}"), index=kind); dt = time.perf_counter() - t0
            print("open", n, kind, db.index_kind, round(dt, 3)); db.close()
