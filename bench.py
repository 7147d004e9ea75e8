#!/usr/bin/env python
"""bench.py -- PloidyFrost per-superbubble hot path on B200 (see DESIGN.md "Measurement").

A "step" is one pass of the hot path over one batch of synthetic superbubbles of a BASELINE.json config: lookup phase A
(CDBG::readCov of the entrance unitig and of every branch), SeqAlign::SequenceAlignment of every bubble's branches, lookup
phase B (site k-mers of the branching bubbles, pf_site_cov).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]      our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                                   the reference's own CPU code (oracle/_ref)

--config 2 (default, the headline): BASELINE configs[2], synthetic hexaploid 1 Gbp -- 3 subgenomes x 333 Mbp, 10 % diverged,
             2 haplotypes each, 90x-equivalent KMC2 database (~1.2 G distinct 25-mers, 2^30-bucket hash index, 34 GB)
--config 1: configs[1], synthetic tetraploid 100 Mbp, 60x (the round-1 headline)
--config 4: configs[4], indel-heavy stress: branches log-uniform 50 bp .. 5 kbp, >= 1 long indel per bubble

`value` = bubbles/s with the batch resident in HBM (CUDA events, max over ranks); `e2e` = the same through the host-pointer
C ABI (pf_kmc_cov_async | pf_align | pf_site_cov | pf_kmc_wait) from pinned host buffers, copies inside the timed region;
`parity` = the results of the timed batch diffed against the unmodified reference (oracle/_ref) on a sample of that batch --
any mismatch makes the run exit non-zero.  At N > 1 the same run also times the KMC index PARTITIONED across the GPUs
(peer-memory lookups and NCCL all-to-all routing) on the same batch and checks it against the replicated index (`sharded`).
"""
from __future__ import annotations

import argparse
import fcntl
import json
import os
import shutil
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 25
SEED0 = 20261017               # seed = 20261017 + config index
SUB_DIVERGENCE = 0.0527        # per-subgenome substitution rate from the root: 10 % between two subgenomes


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: keep this rank's host threads -- and through first touch its pinned staging memory -- on the NUMA node the
    GPU hangs off (PCI bus id from NVML -> /sys/bus/pci/devices/<id>/numa_node -> that node's cpulist).  Returns a small report, or
    the reason nothing was done (single-node hosts, containers without sysfs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # NVML pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"bound": False, "why": "the device reports no NUMA node"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": f"no allowed CPU on node {node}"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus), "pci": bus}
    except Exception as e:   # noqa: BLE001
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU during the timed region, through NVML in-process (an `nvidia-smi`
    subprocess per sample initialises NVML for every GPU of the box each time and perturbs a multi-rank run)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []       # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.stop_flag = False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        while not self.stop_flag and self.h is not None:
            try:
                mhz = int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    rs = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for n, bit in self.REASONS.items() if any(s[1] & bit for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples),
                "source": "NVML (nvmlDeviceGetClockInfo / CurrentClocksEventReasons), 10 ms period, timed region + e2e loop"}


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
class Config:
    """One BASELINE.json config: how its genome is made of subgenomes / haplotypes, and the defaults that go with it."""

    def __init__(self, idx, genome_mbp):
        self.idx = idx
        self.seed = SEED0 + idx
        if idx == 2:      # hexaploid: 3 subgenomes, 2 haplotypes each; 90x over 6 haplotype copies
            self.genome_mbp = genome_mbp or 1000.0
            self.n_sub, self.n_hap = 3, 2
            self.batch, self.ref_sample, self.cpu_sample = 1048576, 262144, 524288
            self.bubbles_per_mbp = 12600
            self.name = (f"configs[2]: synthetic hexaploid {self.genome_mbp:g} Mbp (3 subgenomes x {self.genome_mbp / 3:.4g} Mbp, 10% diverged, "
                         f"2 haplotypes each, 1% SNP, 0.1% indel), 90x-equivalent KMC2 db (k=25, p=9, sig 9, 512 bins), -z 8, M/D/G = 2/-1/-3")
        elif idx == 1:    # tetraploid: one genome, 4 haplotypes; 60x over 4 copies
            self.genome_mbp = genome_mbp or 100.0
            self.n_sub, self.n_hap = 1, 4
            self.batch, self.ref_sample, self.cpu_sample = 262144, 65536, 131072
            self.bubbles_per_mbp = 14000
            self.name = (f"configs[1]: synthetic tetraploid {self.genome_mbp:g} Mbp (4 haplotypes, 1% SNP, 0.1% indel), "
                         f"60x-equivalent KMC2 db (k=25, p=9, sig 9, 512 bins), -z 8, M/D/G = 2/-1/-3")
        elif idx == 4:    # indel-heavy stress
            self.genome_mbp = 0.0
            self.n_sub, self.n_hap = 0, 0
            self.batch, self.ref_sample, self.cpu_sample = 2048, 32, 48
            self.bubbles_per_mbp = 0
            self.name = ("configs[4]: indel-heavy stress, 2-4 branches per bubble, branch lengths log-uniform 50 bp .. 5 kbp, >= 1 long "
                         "indel (5-40 % of the branch) per branch pair, KMC2 db of the batch's own k-mers, M/D/G = 2/-1/-3")
        else:
            raise SystemExit(f"bench.py: --config {idx} is not a bench workload (1, 2 or 4)")
        self.lam = 15.0 * 126.0 / 150.0    # 15x per haplotype copy; k-mer coverage = depth * (L - k + 1) / L
        self.sub_len = int(self.genome_mbp * 1e6 / max(self.n_sub, 1))

    def sub_seed(self, s):
        return self.seed * 1000003 + s

    def root_seed(self):
        return (self.seed * 7919 + 17) if self.n_sub > 1 else 0


def cache_dir(need_bytes):
    """Where the synthetic database lives between the arms / runs of one box: RAM disk when it has room, else /tmp."""
    env = os.environ.get("PF_BENCH_CACHE")
    if env:
        os.makedirs(env, exist_ok=True)
        return env
    for base in ("/dev/shm", "/tmp"):
        try:
            if os.path.isdir(base) and shutil.disk_usage(base).free > need_bytes * 1.25 + (2 << 30):
                d = os.path.join(base, "pfbench_cache")
                os.makedirs(d, exist_ok=True)
                return d
        except OSError:
            pass
    d = "/tmp/pfbench_cache"
    os.makedirs(d, exist_ok=True)
    return d


def make_bubbles(cfg, args, region_rank):
    """The batch of rank `region_rank`: bubbles of its own stretch of every subgenome (generated on its own: the synthetic
    genome is keyed by absolute position, so the stretch carries the whole genome's bases and variants)."""
    from ploidyfrost_b200.synth import workload as wl
    nt = min(16, os.cpu_count() or 8)
    if cfg.idx == 4:
        w = wl.Workload(cfg.seed, 1000, 1, n_threads=1)
        bb = w.long_bubbles(cfg.seed * 31 + region_rank, K, args.batch, 50, 5000, 4)
        w.close()
        return bb
    per_sub = (args.batch + cfg.n_sub - 1) // cfg.n_sub
    margin = 6000
    region = int(per_sub / cfg.bubbles_per_mbp * 1e6 * 1.15) + 200000 + 2 * margin
    region = min(region, cfg.sub_len)
    parts = []
    for s in range(cfg.n_sub):
        r0 = (region_rank * region) % max(1, cfg.sub_len - region + 1)
        w = wl.Workload(cfg.sub_seed(s), region, cfg.n_hap, p_snp=0.01, p_indel=0.001, n_threads=nt, root_seed=cfg.root_seed(),
                        divergence=SUB_DIVERGENCE if cfg.n_sub > 1 else 0.0, offset=r0)
        want = per_sub if s + 1 < cfg.n_sub else args.batch - per_sub * (cfg.n_sub - 1)
        parts.append(w.bubbles(K, margin, region - margin, want))
        w.close()
    return wl.BubbleBatch.concat(parts)


def ensure_db(cfg, args, device, bb=None, tag=""):
    """The KMC database of the config: built once per box (GPU sort/unique with torch -- data tooling -- or numpy for small
    cases) under an exclusive lock, then reused by every arm / rank / run.  Returns (prefix, info)."""
    from ploidyfrost_b200.synth import workload as wl
    est = int(cfg.genome_mbp * 1e6 * 1.3 * 6 + (1 << 30) + (1 << 28)) if cfg.idx != 4 else 1 << 30
    d = cache_dir(est)
    name = f"c{cfg.idx}_g{cfg.genome_mbp:g}_s{cfg.seed}{tag}" if cfg.idx != 4 else f"c4_b{args.batch}_s{cfg.seed}{tag}"
    prefix = os.path.join(d, name)
    meta = prefix + ".json"
    with open(prefix + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        if os.path.exists(meta):
            return prefix, json.load(open(meta))
        t0 = time.perf_counter()
        tmp = prefix + ".tmp"
        if cfg.idx == 4:
            seqs = [bb.bases[int(bb.seq_off[i]):int(bb.seq_off[i + 1])] for i in range(bb.n_seq)]
            seqs += [bb.ent_bases[int(bb.ent_off[i]):int(bb.ent_off[i + 1])] for i in range(bb.n_bubbles)]
            info, _, _ = wl.write_db_numpy(tmp, seqs, K, cfg.lam, cfg.seed, version=0x200, lut_prefix_len=9, sig_len=9, n_bins=64)
        else:
            nt = min(32, os.cpu_count() or 8)
            groups = []
            for s in range(cfg.n_sub):
                w = wl.Workload(cfg.sub_seed(s), cfg.sub_len, cfg.n_hap, p_snp=0.01, p_indel=0.001, n_threads=nt, root_seed=cfg.root_seed(),
                                divergence=SUB_DIVERGENCE if cfg.n_sub > 1 else 0.0)
                groups.append([w.haplotype(i) for i in range(cfg.n_hap)])
                w.close()
            if device is not None:
                info = wl.write_db_torch(tmp, groups, K, cfg.lam, cfg.seed, device=device, version=0x200, lut_prefix_len=9, sig_len=9,
                                         n_bins=512)
            else:
                flat = [h for g in groups for h in g]
                info, _, _ = wl.write_db_numpy(tmp, flat, K, cfg.lam, cfg.seed, version=0x200, lut_prefix_len=9, sig_len=9,
                                               n_bins=512 if cfg.genome_mbp > 5 else 64)
            del groups
        info["build_s"] = round(time.perf_counter() - t0, 1)
        os.replace(tmp + ".kmc_pre", prefix + ".kmc_pre")
        os.replace(tmp + ".kmc_suf", prefix + ".kmc_suf")
        json.dump(info, open(meta, "w"))
        return prefix, info


def workload_config(cfg, args, n_bubbles, info, extra=None):
    c = {"workload": cfg.name, "batch_bubbles": int(n_bubbles), "db_kmers": (info or {}).get("N"),
         "step": "lookup-A (readCov of entrance + branch unitigs) + SequenceAlignment per bubble + lookup-B (site k-mers of the branching bubbles)",
         "l2": "flushed between timed steps (256 MiB memset); the KMC index exceeds L2" if cfg.idx != 4 else
               "flushed between timed steps (256 MiB memset); flag matrices of the long pairs exceed L2",
         "kmc_index": "replicated on every GPU"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference (oracle/_ref), all host threads
# ---------------------------------------------------------------------------------------------------------------------
def open_checker():
    from oracle.bindings import Checker
    try:
        return Checker("ref"), "reference"
    except Exception:
        return Checker("oracle"), "port"


def site_kmer_proxy(msa, strict, k):
    """The lookup-B workload of an alignment result as a flat k-mer batch: for every variable column of every branching bubble and
    every row, the k characters ending at the column (the reference's site k-mer when no indel precedes the site,
    CDBG.cpp:2469-2472); windows that touch a gap or start before the row are left out.  Stands in for the reference's own
    string handling in the TIMED reference legs (same number and kind of database lookups)."""
    nv = np.diff(msa["var_off"]).astype(np.int64)
    b_of_site = np.repeat(np.arange(len(nv)), nv)
    keep = strict[b_of_site] == 0
    b_of_site = b_of_site[keep]
    col = msa["var_col"].astype(np.int64)[keep]
    nr = msa["n_rows"].astype(np.int64)[b_of_site]
    site_of_row = np.repeat(np.arange(len(b_of_site)), nr)
    r = np.arange(len(site_of_row)) - np.repeat(np.cumsum(nr) - nr, nr)
    b = b_of_site[site_of_row]
    L = msa["aln_len"].astype(np.int64)[b]
    c = col[site_of_row]
    ok = c - k + 1 >= 0
    start = (msa["rows_off"].astype(np.int64)[b] + r * L + c - k + 1)[ok]
    if len(start) == 0:
        return np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    win = msa["rows"][start[:, None] + np.arange(k)[None, :]]
    win = win[~(win == ord("-")).any(axis=1)]
    return np.ascontiguousarray(win.reshape(-1)), (np.arange(len(win) + 1, dtype=np.uint64) * k)


def reference_arm(args, cfg, rank):
    """`--impl reference`: the reference's own CPU implementation (unmodified SeqAlign + KMC API in oracle/_ref, driven by
    oracle/ref_shim.cpp with PloidyFrost's readCov call pattern), all host threads, bounded sample per step."""
    if rank != 0:
        return
    ref, kind = open_checker()
    cores = os.cpu_count() or 1
    device = None
    try:
        import torch
        if torch.cuda.is_available():
            device = "cuda:0"
    except Exception:
        pass
    bb = make_bubbles(cfg, args, 0)
    prefix, info = ensure_db(cfg, args, device, bb)
    sample = bb.slice(0, min(bb.n_bubbles, args.ref_sample))
    lb, lo = sample.lookup_sequences()
    h = ref.kmc_open(prefix)
    times = []
    n_lookups = int(np.maximum(np.diff(lo).astype(np.int64) - K + 1, 0).sum())
    n_site_kmers = 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ref.kmc_cov(h, lb, lo, mode=1, low=args.low, up=args.up, n_threads=cores)
        msa = ref.align(sample.bases, sample.seq_off, sample.bubble_off, n_threads=cores)
        t1 = time.perf_counter()
        sb, so = site_kmer_proxy(msa, sample.bubble_type, K)       # untimed: stands in for the reference's string handling
        t2 = time.perf_counter()
        if len(so) > 1:
            ref.kmc_counts(h, sb, so, K, mode=1, n_threads=cores)  # lookup-B
        n_site_kmers = len(so) - 1
        dt = (time.perf_counter() - t2) + (t1 - t0)
        if it >= args.warmup:
            times.append(dt)
    ref.kmc_close(h)
    ms = 1e3 * sum(times) / len(times)
    val = sample.n_bubbles / (ms * 1e-3)
    line = {"impl": "reference", "metric": "superbubble variants/sec", "value": val, "unit": "bubbles/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(cfg, args, args.batch, info, {"sample_bubbles_per_step": sample.n_bubbles}),
            "kmc_lookups_per_s": n_lookups / (ms * 1e-3),
            "cpu_baseline": {"value": val, "unit": "bubbles/s", "cores": cores, "kind": kind,
                             "sample": f"{sample.n_bubbles} bubbles ({n_lookups} k-mer lookups + {n_site_kmers} site k-mer lookups) of the same batch per step"},
            "e2e": {"value": val, "unit": "bubbles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def slice_msa(m, nb):
    """first nb bubbles of a pf_msa_batch_t dump"""
    out = {"n_bubbles": nb}
    for key in ("status", "n_rows", "aln_len"):
        out[key] = m[key][:nb]
    for name, arrs in (("rows", ("rows",)), ("var", ("var_col", "var_kind")), ("cls", ("cls",)), ("ilen", ("ilen",))):
        off = m[name + "_off"][:nb + 1]
        out[name + "_off"] = off
        for a in arrs:
            out[a] = m[a][:int(off[-1])]
    return out


def cpu_baseline_and_parity(args, cfg, bb, prefix, n_sample, gpu_cov, gpu_msa, gpu_sites, skip_np, timed=True):
    """oracle/_ref (the unmodified reference) on this box's host cores, on the first n_sample bubbles of the timed batch, same
    database: its time is `cpu_baseline`, its results are what the CUDA path's results of those bubbles are diffed against."""
    from oracle import parity
    ref, kind = open_checker()
    cores = os.cpu_count() or 1
    S = min(bb.n_bubbles, n_sample)
    sample = bb.slice(0, S)
    lb, lo = sample.lookup_sequences()
    n_lookups = int(np.maximum(np.diff(lo).astype(np.int64) - K + 1, 0).sum())
    h = ref.kmc_open(prefix)
    t0 = time.perf_counter()
    cov_ref = ref.kmc_cov(h, lb, lo, mode=1, low=args.low, up=args.up, n_threads=cores)
    t1 = time.perf_counter()
    msa_ref = ref.align(sample.bases, sample.seq_off, sample.bubble_off, n_threads=cores)
    t2 = time.perf_counter()
    t_site = [0.0, 0]

    def lookup(b, o):   # lookup-B of the reference: 'as written, else reverse complement' per site k-mer
        ta = time.perf_counter()
        r = ref.kmc_counts(h, b, o, K, mode=1, use_read_api=False, n_threads=cores)
        t_site[0] += time.perf_counter() - ta
        t_site[1] += len(o) - 1
        return r

    # ---- parity of the timed batch ----
    nseq_s = int(sample.bubble_off[-1])
    # gpu_cov: the records of exactly these sequences (a sub-batch call), or of the whole batch (entrances first, then branches)
    g_cov = gpu_cov if len(gpu_cov) == S + nseq_s else np.concatenate([gpu_cov[:S], gpu_cov[bb.n_bubbles:bb.n_bubbles + nseq_s]])
    cov_bad = parity.compare_cov(g_cov, cov_ref)
    g_msa = slice_msa(gpu_msa, S)
    msa_bad = int(parity.compare_msa(g_msa, msa_ref).sum())
    site_bad, n_checked = 0, 0
    if gpu_sites is not None:
        checked, st, ncls, cv = parity.expected_site_cov(msa_ref, skip_np[:S], K, args.low, args.up, lookup, max_general=args.parity_general_sites)
        ns, nc = int(msa_ref["var_off"][-1]), int(msa_ref["cls_off"][-1])
        g_sites = {"status": gpu_sites["status"][:ns], "n_class": gpu_sites["n_class"][:ns], "cov": gpu_sites["cov"][:nc]}
        site_bad = parity.compare_site_cov(g_sites, msa_ref, checked, st, ncls, cv) if msa_bad == 0 else -1
        n_checked = int(checked.sum())
    ref.kmc_close(h)
    par = {"against": "oracle/_ref (unmodified SeqAlign + CKMCFile)" if kind == "reference" else "oracle port",
           "bubbles": int(S), "cov_records": int(len(cov_ref)), "cov_mismatches": int(cov_bad),
           "msa_bubbles_all_fields": int(S), "msa_mismatches": int(msa_bad),
           "site_columns_checked": n_checked, "site_mismatches": int(site_bad),
           "mismatches": int(cov_bad + msa_bad + max(site_bad, 0) + (1 if site_bad < 0 else 0))}
    total = (t2 - t0) + t_site[0]
    base = {"value": S / total, "unit": "bubbles/s", "cores": cores, "kind": kind,
            "sample": f"{S} bubbles / {n_lookups} k-mer lookups + {t_site[1]} site k-mer lookups of the same batch, one pass, same KMC db",
            "kmc_lookups_per_s": n_lookups / max(t1 - t0, 1e-9), "align_bubbles_per_s": S / max(t2 - t1, 1e-9),
            "site_kmer_lookups_per_s": t_site[1] / max(t_site[0], 1e-9), "seconds": round(total, 2)}
    return (base if timed else None), par


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configs index: 2 hexaploid 1 Gbp (headline), 1 tetraploid 100 Mbp, 4 indel-heavy")
    ap.add_argument("--genome-mbp", type=float, default=None, help="override the genome size of the config (smaller test runs)")
    ap.add_argument("--batch", type=int, default=None, help="bubbles per step per GPU")
    ap.add_argument("--ref-sample", type=int, default=None, help="bubbles per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=None, help="bubbles of the cpu_baseline / parity leg")
    ap.add_argument("--parity-general-sites", type=int, default=4000,
                    help="parity: variable columns behind an indel site checked through the pure-Python restatement (all others are checked vectorised)")
    ap.add_argument("--low", type=int, default=2)
    ap.add_argument("--up", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg AND the parity block (profiling runs)")
    ap.add_argument("--no-sharded-legs", action="store_true", help="N > 1: skip the partitioned-index legs")
    ap.add_argument("--lookup-sms", type=int, default=0,
                    help="SMs of the partition the lookups run on beside the alignment (pf_lookup_partition); 0 = phases in sequence on one stream")
    ap.add_argument("--lookup-sms-sweep", default="", help="diagnostics: also time the device step with these partition sizes, e.g. 0,16,24,32,48")
    ap.add_argument("--e2e-threads", type=int, default=4, help="host threads of the e2e leg (one pf_ctx + shared index handle each)")
    ap.add_argument("--no-staged-align", action="store_true",
                    help="e2e leg: send the branches a second time with pf_align instead of aligning the copy the lookup call staged (pf_align_staged)")
    ap.add_argument("--e2e-sweep", default="", help="also time the e2e leg with these host thread counts (T) or T x sub-batches per thread (TxC), e.g. 1,2,4x4")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not bind the rank to the NUMA node of its GPU (A/B runs)")
    ap.add_argument("--e2e-profile", default="", help="diagnostics: write a CUPTI timeline (chrome trace) of two e2e steps to this path")
    ap.add_argument("--e2e-trace", action="store_true", help="diagnostics: blocking time of every call of a single-threaded, single-batch e2e pass")
    ap.add_argument("--e2e-chunks", type=int, default=1, help="sub-batches per host thread and step in the e2e leg")
    ap.add_argument("--region-rank", type=int, default=None,
                    help="diagnostics: take the batch another rank would take (its region of the genome) on this GPU")
    ap.add_argument("--no-peer-lookup", action="store_true",
                    help="with --sharded-db: keep the route / NCCL all-to-all / scatter path instead of mapping the other partitions "
                         "through CUDA IPC and loading their buckets over NVLink")
    ap.add_argument("--sharded-db", action="store_true",
                    help="ONLY the partitioned KMC index (no replicated copy): for databases beyond one GPU's HBM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = Config(args.config, args.genome_mbp)
    args.batch = args.batch or cfg.batch
    args.ref_sample = args.ref_sample or cfg.ref_sample
    args.cpu_sample = args.cpu_sample or cfg.cpu_sample
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    from ploidyfrost_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else {"bound": False, "why": "single process"}
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    t_setup = time.perf_counter()
    bb = make_bubbles(cfg, args, rank if args.region_rank is None else args.region_rank)
    t_bubbles = time.perf_counter() - t_setup
    prefix = info = None
    if rank == 0 or cfg.idx == 4:
        prefix, info = ensure_db(cfg, args, str(dev), bb, tag=f"_r{rank}" if cfg.idx == 4 and world > 1 else "")
    if world > 1 and cfg.idx != 4:
        obj = [prefix, info]
        dist.broadcast_object_list(obj, src=0)
        prefix, info = obj
    barrier()
    torch.cuda.empty_cache()
    t_open = time.perf_counter()
    ctx = capi.Context(local_rank)
    sharded_only = bool(args.sharded_db and world > 1)
    if sharded_only:
        from ploidyfrost_b200 import sharded
        db = capi.KmcDb(ctx, prefix, part=rank, n_parts=world)
        sh = sharded.ShardedKmcDb(db)
        peer = (not args.no_peer_lookup) and sharded.attach_peers(db, dev)
    else:
        db = capi.KmcDb(ctx, prefix)
        peer = False
    if cfg.idx != 4 and not sharded_only and db.index_kind != "hash":
        raise SystemExit(f"bench.py: the headline workload must run on the hash index, pf_kmc_open kept '{db.index_kind}' "
                         f"(build status {db.build_status:#x}): a silent layout fallback is a failed run")
    barrier()
    t_open = time.perf_counter() - t_open
    t_setup = time.perf_counter() - t_setup

    # ---- device-resident batch ----
    lb, lo = bb.lookup_sequences()
    wo = capi.window_offsets(lo, K)
    n_win = int(wo[-1])
    n_lseq = len(lo) - 1
    pad = (-len(lb)) % 16
    d_lb = torch.from_numpy(np.concatenate([lb, np.zeros(pad, np.uint8)])).to(dev)
    d_lo = torch.from_numpy(lo.astype(np.int64)).to(dev)
    d_wo = torch.from_numpy(wo.astype(np.int64)).to(dev)
    d_cov = torch.empty(n_lseq * 24, dtype=torch.uint8, device=dev)
    d_ab = torch.from_numpy(bb.bases).to(dev)
    d_ao = torch.from_numpy(bb.seq_off.astype(np.int64)).to(dev)
    d_bo = torch.from_numpy(bb.bubble_off.astype(np.int32)).to(dev)
    skip_np = np.ascontiguousarray(bb.bubble_type.astype(np.uint8))        # strict bubbles: class coverage = sum of branch means, no site k-mers
    d_skip = torch.from_numpy(skip_np).to(dev)
    do_sites = (not sharded_only) or peer                                   # lookup phase B needs every k-mer reachable from this GPU
    seq_len = np.diff(bb.seq_off)
    max_len, max_rows = int(seq_len.max()), int(np.diff(bb.bubble_off).max())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # a side stream: the legacy default stream's handle is 0, which the C ABI reads as "use the context's own stream",
    # and torch.cuda.Event only sees the stream it is recorded on -- so kernels and events share this explicit stream
    stream = torch.cuda.Stream(dev)
    sptr = stream.cuda_stream
    assert sptr != 0
    torch.cuda.synchronize()

    # SM partition of the lookups (pf_lookup_partition, a green context): lookup-A runs on its own SMs beside the alignment
    part = {"ptr": None, "stream": None, "sms": 0}

    def set_partition(n_sm):
        if n_sm and n_sm >= 8:
            ptr, granted = ctx.lookup_partition(n_sm)
            part.update(ptr=ptr, stream=torch.cuda.ExternalStream(ptr, device=dev), sms=granted)
        else:
            part.update(ptr=None, stream=None, sms=0)

    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

    def step_device(ev=None, kdb=None, route=None, cov_t=None):
        kdb = kdb or db
        cov_t = cov_t if cov_t is not None else d_cov
        side = part["stream"] if (route is None and part["stream"] is not None) else None
        if ev:
            ev[0].record(stream)
        if route is not None:
            _, _, c = route.lookup(d_lb, d_lo, d_wo, n_win, mode=capi.LOOKUP_CANONICAL, low=args.low, up=args.up, stream=stream)
            cov_t[:n_lseq * 24].copy_(c[:n_lseq * 24])
        elif side is not None:      # enqueued FIRST, on its partition: the alignment kernels that follow fill the other SMs
            ev_fork.record(stream)
            side.wait_event(ev_fork)
            if ev:
                ev[4].record(side)
            kdb.lookup_dev(d_lb.data_ptr(), len(lb), d_lo.data_ptr(), d_wo.data_ptr(), n_lseq, n_win, capi.LOOKUP_CANONICAL, args.low,
                           args.up, None, None, cov_t.data_ptr(), part["ptr"])
            if ev:
                ev[5].record(side)
            ev_join.record(side)
        else:
            kdb.lookup_dev(d_lb.data_ptr(), len(lb), d_lo.data_ptr(), d_wo.data_ptr(), n_lseq, n_win, capi.LOOKUP_CANONICAL, args.low,
                           args.up, None, None, cov_t.data_ptr(), sptr)
        if ev:
            ev[1].record(stream)
        ctx.align_dev(d_ab.data_ptr(), len(bb.bases), d_ao.data_ptr(), bb.n_seq, d_bo.data_ptr(), bb.n_bubbles, max_len, max_rows,
                      stream=sptr)
        if ev:
            ev[2].record(stream)
        if route is None and (kdb is not db or do_sites):
            kdb.site_cov_dev(args.low, args.up, d_skip.data_ptr(), sptr)
        if side is not None:
            stream.wait_event(ev_join)
        if ev:
            ev[3].record(stream)

    def timed_loop(**kw):
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
        with torch.cuda.stream(stream):
            for it in range(args.steps):
                flush.fill_(it & 0xFF)
                step_device(evs[it], **kw)
        torch.cuda.synchronize()
        seg = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in evs) / len(evs)
        concurrent = part["stream"] is not None and kw.get("route") is None
        return seg(0, 3), (seg(4, 5) if concurrent else seg(0, 1)), seg(1, 2), seg(2, 3)

    main_route = sh if (sharded_only and not peer) else None
    if args.lookup_sms and main_route is None:
        os.environ["PF_LOOKUP_SMS"] = str(args.lookup_sms)      # the host-pointer calls (pf_kmc_cov_async) use the same partition
        set_partition(args.lookup_sms)
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 1)):
            step_device(route=main_route)
    torch.cuda.synchronize()
    cells = ctx.last_cells
    retry = ctx.last_retry_count
    heavy_q = ctx.last_heavy_queued

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    ms_step, ms_lookup, ms_align, ms_site = timed_loop(route=main_route)
    barrier()
    launches = ctx.launches - launches0
    part_sweep = {}
    if args.lookup_sms_sweep and main_route is None:          # diagnostics: the step with other partition sizes (0 = one stream, phases in sequence)
        for n_s in [int(x) for x in args.lookup_sms_sweep.split(",") if x]:
            set_partition(n_s)
            with torch.cuda.stream(stream):
                step_device()
            torch.cuda.synchronize()
            t_ = timed_loop()
            part_sweep[str(part["sms"] if n_s else 0)] = {"ms_step": t_[0], "ms_lookup": t_[1], "ms_align": t_[2], "ms_site": t_[3]}
        set_partition(args.lookup_sms)

    # ---- e2e: host-pointer C ABI from pinned host buffers ----
    # The batch is handed over the way a multi-threaded host hands it over (the reference walks the graph with -t N threads,
    # CDBG.cpp:1929-1945): T host threads, each with its own pf_ctx and a shared handle on the one index (pf_kmc_share), each
    # pushing its share of the step as sub-batches through pf_kmc_cov_async | pf_align | pf_site_cov | pf_kmc_wait.  The copy-in
    # of one thread's sub-batch overlaps the kernels and the copy-out of another's.  T = 1 is the single-threaded caller.
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    h_wo = pinned(wo)
    workers_all = [(ctx, db)]

    def run_chunk(wctx, wdb, ch, keep=False, trace=None):
        t_tr = time.perf_counter()

        def lap(name):           # --e2e-trace: blocking time of every call of the sequence (host clock)
            nonlocal t_tr
            if trace is not None:
                now = time.perf_counter()
                trace[name] = trace.get(name, 0.0) + (now - t_tr) * 1e3
                t_tr = now
        if main_route is not None:   # no host-pointer form of the routed lookup: copy in, route/exchange/lookup, copy the cov records out
            with torch.cuda.stream(stream):
                e_lb = ch["lb"][0].to(dev, non_blocking=True)
                e_lo = ch["lo"][0].to(dev, non_blocking=True).view(torch.int64)
                e_wo = h_wo[0].to(dev, non_blocking=True).view(torch.int64)
                _, _, e_cov = sh.lookup(e_lb, e_lo, e_wo, n_win, mode=capi.LOOKUP_CANONICAL, low=args.low, up=args.up, stream=stream)
                ch["cov_t"].copy_(e_cov[:n_lseq * 24], non_blocking=True)
            stream.synchronize()
            cv = ch["cov"]
        else:   # lookup-A runs on the handle's own stream beside the alignment (pf_kmc_cov_async ... pf_kmc_wait)
            cv = wdb.cov_async(ch["lb"][1], ch["lo"][1], ch["cov"], mode=capi.LOOKUP_CANONICAL, low=args.low, up=args.up)
        lap("pf_kmc_cov_async (enqueue)")
        if main_route is None and not args.no_staged_align:   # the branches are in the lookup batch already: aligned where they were staged
            m = wctx.align_staged(wdb, ch["n"], ch["bo"][1], ch["max_len"], ch["max_rows"], copy=keep)
        else:
            m = wctx.align(ch["ab"][1], ch["ao"][1], ch["bo"][1], copy=keep)   # copy=False: views of the pinned result arena (C-ABI ownership rule)
        lap("pf_align")
        st = wdb.site_cov(args.low, args.up, ch["skip"], copy=keep) if do_sites else None
        lap("pf_site_cov")
        if main_route is None:
            wdb.wait()
        lap("pf_kmc_wait")
        # what crossed the bus: every array of the result except the four offset arrays, which the library rebuilds on the host
        # from the per-bubble counts it copies instead (2 x 4 bytes per bubble)
        nb = cv.nbytes + sum(v.nbytes for k_, v in m.items() if isinstance(v, np.ndarray) and k_ not in ("rows_off", "var_off", "cls_off", "ilen_off")) + 8 * ch["n"]
        if st is not None:   # site_off / cov_off are views of the alignment's var_off / cls_off (counted above): they do not cross the bus again
            nb += sum(v.nbytes for k_, v in st.items() if k_ not in ("site_off", "cov_off"))
        return cv, m, st, nb

    def e2e_leg(T_req, C_req=None, profile_path=None):
        """-> (ms per step, h2d bytes, d2h bytes, threads, sub-batches) of the e2e leg with T_req host threads, C_req sub-batches each"""
        T_e2e = 1 if main_route is not None else max(1, T_req)
        n_chunks = T_e2e * max(1, C_req or args.e2e_chunks) if (T_e2e > 1 or C_req) and main_route is None else 1
        chunks = []
        for c in range(n_chunks):
            b0, b1 = bb.n_bubbles * c // n_chunks, bb.n_bubbles * (c + 1) // n_chunks
            sub = bb if n_chunks == 1 else bb.slice(b0, b1)
            clb, clo = (lb, lo) if n_chunks == 1 else sub.lookup_sequences()
            ch = {"lb": pinned(clb), "lo": pinned(clo), "ab": pinned(sub.bases), "ao": pinned(sub.seq_off), "bo": pinned(sub.bubble_off),
                  "skip": np.ascontiguousarray(sub.bubble_type.astype(np.uint8)), "n": sub.n_bubbles}
            cp = torch.empty((len(clo) - 1) * 24, dtype=torch.uint8).pin_memory()
            ch["cov_t"], ch["cov"] = cp, cp.numpy().view(capi.COV_DTYPE)
            ch["max_len"], ch["max_rows"] = int(np.diff(sub.seq_off).max()), int(np.diff(sub.bubble_off).max())
            staged = main_route is None and not args.no_staged_align
            ch["h2d"] = (clb.nbytes + clo.nbytes + (0 if staged else sub.bases.nbytes + sub.seq_off.nbytes) + sub.bubble_off.nbytes +
                         (ch["skip"].nbytes if do_sites else 0))
            chunks.append(ch)
        while len(workers_all) < T_e2e:
            c2 = capi.Context(local_rank)
            workers_all.append((c2, capi.KmcDb(c2, prefix, share_of=db)))
        workers = workers_all[:T_e2e]
        d2h_acc = [0] * T_e2e

        def worker_steps(t, n_steps):
            wctx, wdb = workers[t]
            tot = 0
            for _ in range(n_steps):
                for c in range(t, n_chunks, T_e2e):
                    tot += run_chunk(wctx, wdb, chunks[c])[3]
            d2h_acc[t] = tot // max(n_steps, 1)

        def run_steps(n_steps):
            if T_e2e == 1:
                worker_steps(0, n_steps)
            else:
                th = [threading.Thread(target=worker_steps, args=(t, n_steps)) for t in range(T_e2e)]
                for x_ in th:
                    x_.start()
                for x_ in th:
                    x_.join()
            torch.cuda.synchronize()

        # K steps inside ONE bracket (barrier + synchronize on both sides), as the device-timed region: every host thread pushes
        # its share of every step -- copy-in, calls, copy-out -- back to back, the way a host that walks a graph hands over batch
        # after batch (integration/: block b + 1 is collected and sent while block b is on the device).  The figure of a single
        # step synchronised on its own (head: the first copy-in, tail: the last copy-out, both exposed) is kept beside it.
        run_steps(2)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_steps(args.steps)
        e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        iso = []
        for it in range(max(2, min(args.steps, 5))):
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run_steps(1)
            iso.append(1e3 * (time.perf_counter() - t0))
        e2e_isolated[(T_e2e, n_chunks)] = sum(iso) / len(iso)
        if profile_path:      # diagnostics: a CUPTI timeline (kernels + copies of every stream) of three more steps
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                run_steps(3)
            prof.export_chrome_trace(profile_path)
        return e2e_ms, sum(ch["h2d"] for ch in chunks), sum(d2h_acc), T_e2e, n_chunks, chunks

    e2e_isolated = {}
    e2e_trace = None
    if args.e2e_trace and main_route is None:      # one host thread, whole batch in one sequence of calls, every call timed
        _, _, _, _, _, tr_chunks = e2e_leg(1)
        e2e_trace = {}
        for it in range(3):
            torch.cuda.synchronize()
            tr = {}
            t0_ = time.perf_counter()
            run_chunk(ctx, db, tr_chunks[0], trace=tr)
            torch.cuda.synchronize()
            tr["total"] = (time.perf_counter() - t0_) * 1e3
            e2e_trace = tr
    e2e_sweep = {}
    for spec in [x for x in args.e2e_sweep.split(",") if x]:       # "T" or "TxC": host threads x sub-batches per thread
        T_s, _, C_s = spec.partition("x")
        e2e_sweep[spec] = e2e_leg(int(T_s), int(C_s) if C_s else None)[0]
    ms_e2e, h2d_bytes, d2h_bytes, T_e2e, n_chunks, chunks = e2e_leg(args.e2e_threads, profile_path=args.e2e_profile or None)
    # the results the parity block diffs: one more (untimed) pass of the first bubbles of the batch through the same calls, copied out
    n_keep = min(bb.n_bubbles, args.cpu_sample if world == 1 else min(args.cpu_sample, 32768))
    if n_chunks == 1 and n_keep == bb.n_bubbles:
        cov, msa, sites, _ = run_chunk(ctx, db, chunks[0], keep=True)
        cov = cov.copy()
    else:
        sub = bb.slice(0, n_keep)
        klb, klo = sub.lookup_sequences()
        kch = {"lb": pinned(klb), "lo": pinned(klo), "ab": pinned(sub.bases), "ao": pinned(sub.seq_off), "bo": pinned(sub.bubble_off),
               "skip": np.ascontiguousarray(sub.bubble_type.astype(np.uint8)), "n": sub.n_bubbles,
               "max_len": int(np.diff(sub.seq_off).max()), "max_rows": int(np.diff(sub.bubble_off).max())}
        kp = torch.empty((len(klo) - 1) * 24, dtype=torch.uint8).pin_memory()
        kch["cov_t"], kch["cov"] = kp, kp.numpy().view(capi.COV_DTYPE)
        if main_route is None:
            cov, msa, sites, _ = run_chunk(ctx, db, kch, keep=True)
            cov = cov.copy()
        else:
            cov, msa, sites = None, None, None
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # whole-batch status / site histograms: one untimed single call over the full batch (views)
    if main_route is not None:
        fm = fs = None
    elif n_chunks == 1:
        _, fm, fs, _ = run_chunk(ctx, db, chunks[0])
    else:
        _, fm, fs, _ = run_chunk(ctx, db, {"lb": pinned(lb), "lo": pinned(lo), "ab": pinned(bb.bases), "ao": pinned(bb.seq_off), "bo": pinned(bb.bubble_off),
                                           "skip": skip_np, "n": bb.n_bubbles, "max_len": max_len, "max_rows": max_rows, "cov": torch.empty(n_lseq * 24, dtype=torch.uint8).pin_memory().numpy().view(capi.COV_DTYPE)})
    site_hist = np.bincount(fs["status"], minlength=5).tolist() if fs is not None else None
    n_sites = int(len(fs["status"])) if fs is not None else 0
    n_ok = int((fm["status"] == 0).sum()) if fm is not None else -1
    status_hist = {int(k): int(v) for k, v in zip(*np.unique(fm["status"], return_counts=True))} if fm is not None else None
    full_sites = {k_: v.copy() for k_, v in fs.items()} if fs is not None else None

    # ---- N > 1: the same batch through the KMC index PARTITIONED across the GPUs, checked against the replicated index ----
    shard_res = None
    if world > 1 and not sharded_only and not args.no_sharded_legs and cfg.idx != 4:
        from ploidyfrost_b200 import sharded
        shard_res = {"n_parts": world, "partition": "hash index sliced by mix(key) % n_gpus"}
        rep_cov = d_cov.clone()
        rep_sites = full_sites
        t0 = time.perf_counter()
        dbp = capi.KmcDb(ctx, prefix, part=rank, n_parts=world)
        barrier()
        shard_res["open_s"] = round(time.perf_counter() - t0, 2)
        shard_res["index_bytes_per_gpu"] = dbp.device_bytes
        shard_res["index_kind"] = dbp.index_kind
        d_cov2 = torch.empty_like(d_cov)
        shp = sharded.ShardedKmcDb(dbp)
        with torch.cuda.stream(stream):        # (a) route -> NCCL all-to-all -> lookup at the owner -> all-to-all -> scatter
            step_device(kdb=dbp, route=shp, cov_t=d_cov2)
        torch.cuda.synchronize()
        barrier()
        r_step, r_lookup, _, _ = timed_loop(kdb=dbp, route=shp, cov_t=d_cov2)
        eq_route = bool(torch.equal(d_cov2, rep_cov))
        shard_res["nccl_route"] = {"ms_lookup": r_lookup, "equal_to_replicated": eq_route, "keys_sent_per_step": shp.last_sent}
        barrier()
        eq_peer = eq_sites = None
        if sharded.attach_peers(dbp, dev):     # (b) the other slices mapped through CUDA IPC: one kernel, buckets loaded over NVLink
            d_cov2.zero_()
            with torch.cuda.stream(stream):
                step_device(kdb=dbp, cov_t=d_cov2)
            torch.cuda.synchronize()
            barrier()
            p_step, p_lookup, p_align, p_site = timed_loop(kdb=dbp, cov_t=d_cov2)
            eq_peer = bool(torch.equal(d_cov2, rep_cov))
            psites = dbp.site_cov(args.low, args.up, skip_np, copy=False)
            eq_sites = all(np.array_equal(psites[k_], rep_sites[k_]) for k_ in rep_sites)
            shard_res["peer_memory"] = {"ms_step": p_step, "ms_lookup": p_lookup, "ms_site_cov": p_site, "equal_to_replicated": eq_peer,
                                        "site_cov_equal_to_replicated": bool(eq_sites)}
        else:
            shard_res["peer_memory"] = None
        barrier()
        flags = torch.tensor([int(eq_route), int(eq_peer is not False), int(eq_sites is not False)], dtype=torch.int32, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        shard_res["all_ranks_equal"] = bool(flags.min().item())
        tt = torch.tensor([r_lookup, shard_res["peer_memory"]["ms_lookup"] if shard_res["peer_memory"] else 0.0,
                           shard_res["peer_memory"]["ms_step"] if shard_res["peer_memory"] else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        shard_res["nccl_route"]["ms_lookup"] = float(tt[0])
        if shard_res["peer_memory"]:
            shard_res["peer_memory"]["ms_lookup"] = float(tt[1])
            shard_res["peer_memory"]["ms_step"] = float(tt[2])
        dbp.close()

    # ---- reduce over ranks (max time, summed work) ----
    sys.stderr.write(f"[rank {rank}] step {ms_step:.2f} ms (lookup {ms_lookup:.2f}, align {ms_align:.2f}, sites {ms_site:.2f}), "
                     f"e2e {ms_e2e:.2f} ms (numa {numa}); tiers {ctx.last_tier_counts} heavy-queued {heavy_q}; setup {t_setup:.1f}s "
                     f"(bubbles {t_bubbles:.1f}s, open {t_open:.1f}s); batch {bb.stats()}\n")
    from ploidyfrost_b200 import shard
    (ms_step, ms_e2e, ms_lookup, ms_align, ms_site), (tot_bubbles, tot_win, tot_cells) = shard.reduce_step(
        [ms_step, ms_e2e, ms_lookup, ms_align, ms_site], [bb.n_bubbles, n_win, cells], device=dev)

    rc = 0
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        gather_gbs = ctx.bench_random_gather(4 << 30)
        int32_gops = ctx.bench_int32()
        # DRAM traffic per launch comes from an ncu capture of THIS workload (profiles/r02_traffic.json); it is not measured in the run
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(f"config{cfg.idx}", {})
        except Exception:
            pass
        default_wl = args.batch == cfg.batch and args.genome_mbp is None and not sharded_only
        tr_lookup = traffic.get("kmc_hash_lookup_kernel") if default_wl and db.index_kind == "hash" else None
        tr_align = traffic.get("align_pipeline") if default_wl else None
        lookups_s = (tot_win / world) / (ms_lookup * 1e-3)       # per GPU, for the per-kernel roofline
        cells_s = (tot_cells / world) / (ms_align * 1e-3)
        roof_lookup = {"kernel": "kmc_hash_lookup_kernel" if db.index_kind == "hash" else "kmc_lookup_kernel", "bound": "hbm",
                       "achieved": lookups_s * 64 / 1e9, "peak": hbm_peak,
                       "unit": "GB/s", "frac": lookups_s * 64 / 1e9 / hbm_peak,
                       "traffic": (tr_lookup or {}).get("bytes"), "traffic_source": (tr_lookup or {}).get("source"),
                       "peak_source": peak_src,
                       "algorithmic_bytes_per_lookup": 64, "lookups_per_launch": tot_win / world, "ms": ms_lookup,
                       "random_sector_gather_gbs": gather_gbs, "index": db.index_kind, "index_bytes": db.device_bytes,
                       "sectors_per_lookup_design": 1,
                       "frac_of_random_sector_rate": lookups_s * 32 / 1e9 / gather_gbs}
        roof_align = {"kernel": "alignment pipeline (msa_lane_kernel / msa_group_kernel / msa_cta_kernel by size class, concurrent streams)",
                      "bound": "int32", "achieved": cells_s * 18 / 1e9, "peak": int32_gops,
                      "unit": "Gop/s", "frac": cells_s * 18 / 1e9 / int32_gops,
                      "traffic": (tr_align or {}).get("bytes"), "traffic_source": (tr_align or {}).get("source"),
                      "peak_source": "measured in this run (pf_bench_int32: IADD/IMNMX/LOP mix); MEASURED_PEAKS.json has no INT32 entry",
                      "int32_ops_per_cell": 18, "cells_per_launch": tot_cells / world, "ms": ms_align}
        dominant = roof_align if ms_align >= ms_lookup else roof_lookup
        extra = {}
        if sharded_only:
            extra["kmc_index"] = ("hash index partitioned by mix(key) % n_gpus; " +
                                  ("other partitions mapped through CUDA IPC, buckets loaded over NVLink inside the lookup kernel" if peer
                                   else "queries routed by NCCL all-to-all"))
        line = {"metric": "superbubble variants/sec", "value": tot_bubbles / (ms_step * 1e-3), "unit": "bubbles/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": workload_config(cfg, args, bb.n_bubbles, info, extra),
                "kmc_lookups": {"metric": "KMC k-mer lookups/sec", "value": tot_win / (ms_lookup * 1e-3), "unit": "lookups/s",
                                "lookups_per_step": tot_win, "ms_lookup_kernel": ms_lookup,
                                "per_step_rate": tot_win / (ms_step * 1e-3)},
                "dp_cells_per_s_kernel": tot_cells / (ms_align * 1e-3),
                "ms_lookup_kernel": ms_lookup, "ms_align_pipeline": ms_align, "ms_site_cov_kernel": ms_site,
                "phase_overlap": ({"lookup_sm_partition": part["sms"], "how": "lookup-A on a green-context stream confined to that many SMs, enqueued before the alignment "
                                   "pipeline of the same step; ms_lookup_kernel / ms_align_pipeline are their own (overlapping) durations"}
                                  if part["stream"] is not None else {"lookup_sm_partition": 0, "how": "phases in sequence on one stream"}),
                "sm_partition_sweep": part_sweep or None,
                "site_columns_per_step": n_sites, "site_status_hist(ok,dropped,missing,undefined,skipped)": site_hist,
                "clocks": sampler.summary(),
                "e2e": {"value": tot_bubbles / (ms_e2e * 1e-3), "unit": "bubbles/s", "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e,
                        "host_threads": T_e2e, "sub_batches_per_step": n_chunks,
                        "ms_per_step_by_host_threads": {str(k_): v for k_, v in e2e_sweep.items()} or None,
                        "single_thread_call_ms": e2e_trace,
                        "numa": numa,
                        "timed": f"{args.steps} steps in one bracket (barrier + synchronize on both sides), host threads free-running",
                        "ms_per_step_synchronised_alone": e2e_isolated.get((T_e2e, n_chunks)),
                        "calls": ("pf_kmc_cov_async | " + ("pf_align" if (args.no_staged_align or main_route is not None) else "pf_align_staged") + " | pf_site_cov | pf_kmc_wait per sub-batch; one pf_ctx + pf_kmc_share handle per host thread")},
                "gpu_launches": int(launches),
                "roofline": dominant, "roofline_lookup": roof_lookup, "roofline_align": roof_align,
                "bubbles_ok": n_ok, "bubble_status_hist": status_hist, "tier2_retries": int(retry), "heavy_queued": int(heavy_q),
                "setup_s": round(t_setup, 1), "db_open_s": round(t_open, 2), "db_build_s": (info or {}).get("build_s"),
                "batch_stats": bb.stats()}
        if shard_res is not None:
            line["sharded"] = shard_res
            if not shard_res.get("all_ranks_equal", True):
                rc = 3
        if not args.no_cpu_baseline and cov is not None:
            base, par = cpu_baseline_and_parity(args, cfg, bb, prefix, n_keep, cov, msa, sites, skip_np, timed=(world == 1))
            if base is not None:
                # the whole reference PROGRAM against the same binary with its estimation phase bound to this library
                # (integration/time_program.py on a real Bifrost graph): captured separately, cited here, not measured in this run
                wp_path = os.path.join(ROOT, "profiles", "r02_whole_program.json")
                if os.path.exists(wp_path):
                    try:
                        base["whole_program"] = json.load(open(wp_path))
                    except ValueError:
                        pass
                line["cpu_baseline"] = base
            line["parity"] = par
            if par["mismatches"]:
                rc = 2
        else:
            line["parity"] = None
        print(json.dumps(line), flush=True)
        if rc:
            sys.stderr.write(f"bench.py: PARITY FAILURE (rc {rc}): {json.dumps(line.get('parity'))} {json.dumps(line.get('sharded'))}\n")
    for wc, wd in workers_all[1:]:
        wd.close()
        wc.close()
    db.close()
    ctx.close()
    if world > 1:
        code = torch.tensor([rc], dtype=torch.int32, device=dev)
        dist.all_reduce(code, op=dist.ReduceOp.MAX)
        rc = int(code.item())
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    main()
