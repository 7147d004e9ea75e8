"""Whole-program timing at scale: the unmodified reference (`oracle/_ref/PloidyFrost -t N`) against the same binary with its
per-superbubble analysis bound to libpfgpu.so (`oracle/_ref/PloidyFrost_gpu`, integration/Makefile) on a REAL Bifrost graph of a
synthetic polyploid genome (BASELINE configs[0] / [1] shapes: 1 % SNP, 0.1 % indel, k = 25).

Both binaries load the same graph and find the same superbubbles with the reference's own code; only the estimation phase
differs.  The graph is built by the reference's own Bifrost from the haplotype sequences (`Bifrost build -r`: the error-free
read graph at full coverage without simulating 10^8 reads), the KMC database holds the canonical k-mers of the haplotypes with
Poisson(depth) counters (ploidyfrost_b200/synth, GPU sort/unique with torch -- data tooling).

Reported per run: wall clock of the whole process, the phase's own `Cpu time` / `Real time` lines, the phase time of the GPU
path, bubbles called; `-t 1` GPU files are compared byte for byte with the reference's `-t 1` files when PF_PROGRAM_CHECK=1
(costs one more reference run), `-t N` files as schedule-independent multisets (tests/e2e_rows.thread_dialect_view).

Usage: python integration/time_program.py GENOME_BP HAPLOTYPES [out.json]        (needs a GPU; run under gpurun)
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import refrun as e2e_rows  # noqa: E402  (reference binaries + multiset views; no checker is loaded here)


def run(binary, cwd, threads, extra_env=None, colored=False):
    env = dict(os.environ)
    env.update(extra_env or {})
    t0 = time.perf_counter()
    # the reference's own phase timer has 1 s resolution (time(NULL)); its messages end with endl, so the arrival times of the
    # phase's first and last line on the pipe give the phase's wall time to a millisecond
    cmd = ([binary, "-g", "dbg.gfa", "-f", "dbg.bfg_colors", "-d", "dbs.txt", "-C", "cov.txt", "-t", str(threads), "-o", "P"] if colored else
           [binary, "-g", "dbg.gfa", "-d", "db", "-t", str(threads), "-l", "2", "-u", "1000", "-o", "P"])
    pr = subprocess.Popen(cmd, cwd=cwd,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    lines, t_begin, t_end = [], None, None
    for ln in pr.stdout:
        now = time.perf_counter()
        lines.append(ln)
        if "Analyzing superbubbles to generate" in ln:
            t_begin = now
        elif "PloidyEstimation():" in ln and "Cpu time" in ln:
            t_end = now
    err = pr.stderr.read()
    pr.wait()

    class R:
        pass
    r = R()
    r.returncode, r.stdout, r.stderr = pr.returncode, "".join(lines), err
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        return {"wall_s": round(wall, 3), "rc": r.returncode, "tail": r.stdout[-600:] + r.stderr[-300:]}
    out = {"wall_s": round(wall, 3)}
    if t_begin is not None and t_end is not None:
        out["estimation_phase_s"] = round(t_end - t_begin, 4)
    m = re.search(r"PloidyEstimation\(\):\s+Cpu time : ([0-9.e+-]+)s", r.stdout)
    if m:
        out["estimation_cpu_s"] = float(m.group(1))
    m = re.search(r"PloidyEstimation\(\):\s+Real time : ([0-9.e+-]+)s", r.stdout)
    if m:
        out["estimation_real_s_1s_resolution"] = float(m.group(1))
    gc = re.search(r"GPU path : (\d+) bubbles, (\d+) colours, (\d+) host threads, phase ([0-9.e+-]+)s = collecting ([0-9.e+-]+)s, waiting for the device ([0-9.e+-]+)s; device thread (.*)", r.stdout)
    if gc:
        out.update(bubbles_walked=int(gc.group(1)), colours=int(gc.group(2)), host_threads=int(gc.group(3)), phase_s=float(gc.group(4)),
                   collect_s=float(gc.group(5)), device_wait_s=float(gc.group(6)), device_thread=gc.group(7))
    g = re.search(r"GPU path : (\d+) bubbles, (\d+) host threads, phase ([0-9.e+-]+)s = waited for device \+ database ([0-9.e+-]+)s, "
                  r"collecting ([0-9.e+-]+)s(?: \([^)]*\))?, waiting for the device ([0-9.e+-]+)s", r.stdout)
    if g:
        out.update(bubbles_walked=int(g.group(1)), host_threads=int(g.group(2)), phase_s=float(g.group(3)), open_wait_s=float(g.group(4)),
                   collect_s=float(g.group(5)), device_wait_s=float(g.group(6)))
    g = re.search(r"releasing the device ([0-9.e+-]+)s", r.stdout)
    if g:
        out["release_s"] = float(g.group(1))
    # every section of the reference prints "<name>: Real time"; keep them all (1 s resolution) for the phase split
    m = re.search(r"GPU path, device thread : (.*)", r.stdout)
    if m:
        out["device_thread"] = m.group(1)
    out["sections"] = re.findall(r"([A-Za-z:()_ ]+?):?\s+Real time : ([0-9.e+-]+)s", r.stdout)
    return out


def main_colored(genome, n_hap, n_samples, out_json):
    """BASELINE configs[3] shape through the real programs: n_samples samples over haplotype subsets of one polyploid, `Bifrost build -c`
    (one colour per sample), one KMC database per sample, `-C` thresholds; unmodified reference against the bound binary, `-t cores`."""
    from ploidyfrost_b200.synth import workload as wl
    pf, bf = e2e_rows.reference_binaries()
    gpu = os.environ.get("PF_PROGRAM_BINARY") or os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_gpu")
    dev = os.environ.get("PF_PROGRAM_DEVICE", "cuda:0")
    cores = os.cpu_count()
    k = 25
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 60 * genome * n_samples else None
    with tempfile.TemporaryDirectory(dir=base) as tmp:
        t0 = time.perf_counter()
        w = wl.Workload(20261017 + 3, genome, n_hap, p_snp=0.01, p_indel=0.001, n_threads=min(16, cores))
        haps = [w.haplotype(i) for i in range(n_hap)]
        w.close()
        prefixes, n_kmers = [], 0
        for s in range(n_samples):
            mine = [h for h in range(n_hap) if (h + s) % 3 != 0]
            if len(mine) < 2:
                mine = list(range(n_hap))[:2]
            with open(os.path.join(tmp, f"s{s}.fa"), "wb") as f:
                for i in mine:
                    f.write(b">s%dh%d\n" % (s, i))
                    f.write(haps[i].tobytes())
                    f.write(b"\n")
            info = wl.write_db_torch(os.path.join(tmp, f"db{s}"), [haps[i] for i in mine], k, 12.6, 20261017 + s, device=dev, version=0x200,
                                     lut_prefix_len=9, sig_len=9, n_bins=64)
            n_kmers += info["N"]
            prefixes.append(f"db{s}")
        del haps
        open(os.path.join(tmp, "dbs.txt"), "w").write("".join(p + "\n" for p in prefixes))
        open(os.path.join(tmp, "cov.txt"), "w").write("".join("2\t1000\n" for _ in prefixes))
        t_data = time.perf_counter() - t0
        t0 = time.perf_counter()
        cmd = [bf, "build", "-c", "-k", str(k), "-i", "-d", "-o", "dbg", "-t", str(min(cores, 16))]
        for s in range(n_samples):
            cmd += ["-r", f"s{s}.fa"]
        subprocess.run(cmd, cwd=tmp, check=True, capture_output=True)
        t_graph = time.perf_counter() - t0
        res = {"genome_bp": genome, "haplotypes": n_hap, "samples": n_samples, "k": k, "db_kmers_all_samples": n_kmers, "host_cores": cores,
               "data_s": round(t_data, 1), "bifrost_build_s": round(t_graph, 1), "runs": {}}

        def fresh(name):
            d = os.path.join(tmp, name)
            shutil.rmtree(d, ignore_errors=True)
            os.mkdir(d)
            for f in os.listdir(tmp):
                if f.startswith("dbg.") or f.startswith("db") or f == "cov.txt":
                    os.symlink(os.path.join(tmp, f), os.path.join(d, f))
            return d

        d_refN, d_gpuN = fresh("refN"), fresh("gpuN")
        res["runs"][f"reference -t {cores}"] = run(pf, d_refN, cores, colored=True)
        res["runs"][f"gpu -t {cores}"] = run(gpu, d_gpuN, cores, colored=True)
        try:
            a = e2e_rows.colored_thread_dialect_view(os.path.join(d_refN, "PloidyFrost_output"))
            b = e2e_rows.colored_thread_dialect_view(os.path.join(d_gpuN, "PloidyFrost_output"))
            res["tN_files_equal_as_multisets"] = bool(a[0] == b[0])
            res["bubbles_called"] = len(b[1])
        except Exception as e:   # noqa: BLE001
            res["tN_files_equal_as_multisets"] = f"not compared: {e}"
        rN, gN = res["runs"][f"reference -t {cores}"], res["runs"][f"gpu -t {cores}"]
        if "estimation_phase_s" in rN and "estimation_phase_s" in gN and gN.get("bubbles_walked"):
            nb = gN["bubbles_walked"]
            res["summary"] = {"workload": f"coloured: {n_samples} samples over {n_hap} haplotypes of {genome / 1e6:g} Mbp, real `Bifrost build -c` graph, {nb} superbubbles walked, "
                                          f"{cores} host cores, -t {cores}",
                              "reference_estimation_phase_s": rN["estimation_phase_s"], "gpu_estimation_phase_s": gN["estimation_phase_s"],
                              "reference_bubbles_per_s": round(nb / rN["estimation_phase_s"]), "gpu_bubbles_per_s": round(nb / gN["estimation_phase_s"]),
                              "estimation_phase_speedup": round(rN["estimation_phase_s"] / gN["estimation_phase_s"], 2),
                              "reference_wall_s": rN["wall_s"], "gpu_wall_s": gN["wall_s"],
                              "measured_by": "integration/time_program.py --colored (wall clock between the phase's first and last console line)"}
    print(json.dumps(res))
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--colored":     # --colored GENOME_BP HAPLOTYPES SAMPLES [out.json]
        return main_colored(int(float(sys.argv[2])), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else None)
    genome = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    n_hap = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    from ploidyfrost_b200.synth import workload as wl
    pf, bf = e2e_rows.reference_binaries()
    gpu = os.path.join(ROOT, "oracle", "_ref", "PloidyFrost_gpu")
    cores = os.cpu_count()
    k = 25
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 40 * genome else None
    with tempfile.TemporaryDirectory(dir=base) as tmp:
        t0 = time.perf_counter()
        w = wl.Workload(20261017, genome, n_hap, p_snp=0.01, p_indel=0.001, n_threads=min(16, cores))
        haps = [w.haplotype(i) for i in range(n_hap)]
        w.close()
        with open(os.path.join(tmp, "haps.fa"), "wb") as f:
            for i, h in enumerate(haps):
                f.write(b">hap%d\n" % i)
                f.write(h.tobytes())
                f.write(b"\n")
        lam = 60.0 / 4 * 126.0 / 150.0 if n_hap == 4 else 30.0 / n_hap * 126.0 / 150.0
        info = wl.write_db_torch(os.path.join(tmp, "db"), haps, k, lam, 20261017, device="cuda:0", version=0x200, lut_prefix_len=9, sig_len=9,
                                 n_bins=512 if genome > 5e6 else 64)
        del haps
        t_data = time.perf_counter() - t0
        t0 = time.perf_counter()
        subprocess.run([bf, "build", "-r", "haps.fa", "-k", str(k), "-i", "-d", "-o", "dbg", "-t", str(min(cores, 16))], cwd=tmp, check=True,
                       capture_output=True)
        t_graph = time.perf_counter() - t0
        res = {"genome_bp": genome, "haplotypes": n_hap, "k": k, "db_kmers": info["N"], "host_cores": cores, "data_s": round(t_data, 1),
               "bifrost_build_s": round(t_graph, 1), "runs": {}}

        def fresh(name):
            d = os.path.join(tmp, name)
            shutil.rmtree(d, ignore_errors=True)
            os.mkdir(d)
            for f in os.listdir(tmp):
                if f.startswith("dbg.") or f.startswith("db.kmc"):
                    os.symlink(os.path.join(tmp, f), os.path.join(d, f))
            return d

        d_refN = fresh("refN")
        if os.environ.get("PF_PROGRAM_SKIP_REF") != "1":
            res["runs"][f"reference -t {cores}"] = run(pf, d_refN, cores)
        d_gpuN = fresh("gpuN")
        res["runs"][f"gpu -t {cores}"] = run(gpu, d_gpuN, cores)
        d_gpu1 = fresh("gpu1")
        res["runs"]["gpu -t 1"] = run(gpu, d_gpu1, 1)
        try:
            a = e2e_rows.thread_dialect_view(os.path.join(d_refN, "PloidyFrost_output"))
            b = e2e_rows.thread_dialect_view(os.path.join(d_gpuN, "PloidyFrost_output"))
            res["tN_files_equal_as_multisets"] = bool(a[0] == b[0])      # [1] = the VarIds in file order: schedule-dependent in the reference
            if a[0] != b[0]:
                # is the reference's own -t N run reproducible?  (its worker threads race on the visited marks, CDBG.cpp:1929-2000)
                d_refN2 = fresh("refN2")
                run(pf, d_refN2, cores)
                a2 = e2e_rows.thread_dialect_view(os.path.join(d_refN2, "PloidyFrost_output"))
                res["reference_tN_equals_its_own_second_run"] = bool(a[0] == a2[0])
                res["tN_differences"] = {k_: [len(a[0][k_]), len(b[0][k_]), len(a2[0][k_])] for k_ in a[0] if a[0][k_] != b[0][k_]}
        except Exception as e:   # noqa: BLE001
            res["tN_files_equal_as_multisets"] = f"not compared: {e}"
        if os.environ.get("PF_PROGRAM_CHECK") == "1":
            import filecmp
            d_ref1 = fresh("ref1")
            res["runs"]["reference -t 1"] = run(pf, d_ref1, 1)
            g1, r1 = os.path.join(d_gpu1, "PloidyFrost_output"), os.path.join(d_ref1, "PloidyFrost_output")
            res["t1_files_identical"] = all(filecmp.cmp(os.path.join(r1, n), os.path.join(g1, n), shallow=False)
                                            for n in os.listdir(r1) if n.startswith("P_") and n != "P_Unitig_Id.txt")
        try:
            res["bubbles_called"] = len({ln.split("\t")[0] for ln in open(os.path.join(d_gpu1, "PloidyFrost_output", "P_alignseq.txt"))})
        except OSError:
            pass
    rN, gN = res["runs"].get(f"reference -t {cores}", {}), res["runs"].get(f"gpu -t {cores}", {})
    if "estimation_phase_s" in rN and "estimation_phase_s" in gN and gN.get("bubbles_walked"):
        nb = gN["bubbles_walked"]
        res["summary"] = {"workload": f"{genome / 1e6:g} Mbp, {n_hap} haplotypes, real Bifrost graph, {nb} superbubbles walked, {cores} host cores, -t {cores}",
                          "reference_estimation_phase_s": rN["estimation_phase_s"], "gpu_estimation_phase_s": gN["estimation_phase_s"],
                          "reference_bubbles_per_s": round(nb / rN["estimation_phase_s"]), "gpu_bubbles_per_s": round(nb / gN["estimation_phase_s"]),
                          "estimation_phase_speedup": round(rN["estimation_phase_s"] / gN["estimation_phase_s"], 2),
                          "reference_wall_s": rN["wall_s"], "gpu_wall_s": gN["wall_s"],
                          "measured_by": "integration/time_program.py (wall clock between the phase's first and last console line)"}
    print(json.dumps(res))
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
