"""Whole-program timing of the reference's driver against the same binary with the per-superbubble analysis bound to libpfgpu.so
(integration/Makefile), BASELINE configs[0]: synthetic diploid 1 Mbp, 30x 150 bp reads, k = 25.  Both binaries load the same graph
and find the same superbubbles with the reference's own code; only the estimation phase differs.  Wall clock of the whole program
and the phase's own `Cpu time` line are reported; the output files of every run are compared (bytes for -t 1).
Usage: [PF_DRIVER_RUNS=ref1,refN,gpu1,gpu1b,gpuN] python integration/time_drivers.py [genome_bp] [out.json]"""
import filecmp
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import e2e_rows  # noqa: E402  (dataset generator shared with the tests)


def run(binary, cwd, threads):
    t0 = time.perf_counter()
    r = subprocess.run([binary, "-g", "dbg.gfa", "-d", "db", "-t", str(threads), "-l", "2", "-u", "1000", "-o", "P"], cwd=cwd,
                       capture_output=True, text=True, check=True)
    wall = time.perf_counter() - t0
    m = re.search(r"PloidyEstimation\(\):\s+Cpu time : ([0-9.e+-]+)s", r.stdout)
    out = {"wall_s": round(wall, 3), "estimation_cpu_s": float(m.group(1)) if m else None}
    g = re.search(r"GPU path : (\d+) bubbles, graph walk ([0-9.e+-]+)s, waited for device \+ database ([0-9.e+-]+)s, lookups \+ alignment \+ rows ([0-9.e+-]+)s",
                  r.stdout)
    if g:
        out.update(bubbles=int(g.group(1)), walk_s=float(g.group(2)), open_s=float(g.group(3)), calls_s=float(g.group(4)))
    return out


def main():
    genome = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    out_json = sys.argv[2] if len(sys.argv) > 2 else None
    pf, _ = e2e_rows.reference_binaries()
    gpu = os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_gpu")
    cores = os.cpu_count()
    with tempfile.TemporaryDirectory() as tmp:
        ref_dir = os.path.join(tmp, "ref")
        os.mkdir(ref_dir)
        e2e_rows.run_reference_config0(ref_dir, genome=genome)       # writes reads, graph, database; runs the reference once (-t 1)
        res = {"genome_bp": genome, "host_cores": cores, "runs": {}}
        golden = os.path.join(tmp, "golden")
        shutil.copytree(os.path.join(ref_dir, "PloidyFrost_output"), golden)
        plan = {"ref1": ("reference -t 1", pf, 1), "refN": (f"reference -t {cores}", pf, cores), "gpu1": ("gpu -t 1", gpu, 1),
                "gpu1b": ("gpu -t 1 (second run)", gpu, 1), "gpuN": (f"gpu -t {cores}", gpu, cores)}
        for key in os.environ.get("PF_DRIVER_RUNS", "ref1,refN,gpu1,gpu1b,gpuN").split(","):
            name, binary, threads = plan[key]
            d = os.path.join(tmp, "run")
            shutil.rmtree(d, ignore_errors=True)
            os.mkdir(d)
            for f in ("dbg.gfa", "db.kmc_pre", "db.kmc_suf"):
                shutil.copy(os.path.join(ref_dir, f), d)
            res["runs"][name] = run(binary, d, threads)
            if threads == 1:
                same = all(filecmp.cmp(os.path.join(golden, n), os.path.join(d, "PloidyFrost_output", n), shallow=False)
                           for n in os.listdir(golden) if n.startswith("P_"))
                res["runs"][name]["files_identical_to_reference_t1"] = same
        n_bubbles = len({ln.split("\t")[0] for ln in open(os.path.join(golden, "P_alignseq.txt"))})
        res["bubbles_called"] = n_bubbles
    print(json.dumps(res))
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
