#!/usr/bin/env python
"""bench.py -- PloidyFrost per-superbubble hot path on B200 (see DESIGN.md "Measurement").

A "step" is one pass of the hot path over one batch of synthetic superbubbles of BASELINE.json configs[1]
(tetraploid, k=25, default scoring): phase-A k-mer coverage lookups (entrance unitig + every branch,
CDBG::readCov) followed by SeqAlign::SequenceAlignment of every bubble's branches.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                           the reference's own CPU code (oracle/_ref)

`value` = bubbles/s with the batch resident in HBM (CUDA events, max over ranks); `e2e` = the same through the
host-pointer C ABI (pf_kmc_cov_async + pf_align + pf_site_cov + pf_kmc_wait) from pinned host buffers, copies inside the
timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 25
SEED = 20261017 + 1            # seed = 20261017 + config index
BUBBLES_PER_MBP = 14000        # measured density of the generator at these rates (only used to size regions)


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


def make_workload(args, rank, world, need_db_files, db_dir, device=None):
    from ploidyfrost_b200.synth import workload as wl
    G = int(args.genome_mbp * 1e6)
    w = wl.Workload(SEED, G, 4, p_snp=0.01, p_indel=0.001, n_threads=min(16, os.cpu_count() or 8))
    region = int(args.batch / BUBBLES_PER_MBP * 1e6 * 1.15) + 200000
    region = min(region, G)
    r0 = (rank * region) % max(1, G - region + 1)
    bb = w.bubbles(K, r0, r0 + region, args.batch)
    prefix = os.path.join(db_dir, "db")
    info = None
    if need_db_files:
        haps = [w.haplotype(i) for i in range(4)]
        lam = 15.0 * 126.0 / 150.0           # 60x over 4 haplotype copies, k-mer coverage = depth*(L-k+1)/L
        if device is not None:
            info = wl.write_db_torch(prefix, haps, K, lam, SEED, device=device, version=0x200, lut_prefix_len=9, sig_len=9,
                                     n_bins=512)
        else:
            info, _, _ = wl.write_db_numpy(prefix, haps, K, lam, SEED, version=0x200, lut_prefix_len=9, sig_len=9,
                                           n_bins=512 if G > 5e6 else 64)
        del haps
    w.close()
    return bb, prefix, info


def site_kmer_proxy(msa, strict, k):
    """CPU legs only: the lookup-B workload of an alignment result as a flat k-mer batch -- for every variable column of every
    branching bubble and every row, the k characters ending at the column (the reference's site k-mer when no indel precedes the
    site, CDBG.cpp:2469-2472); windows that touch a gap or start before the row are left out.  A proxy for the reference's own
    string handling (which cannot be driven without its Bifrost graph): same number and kind of database lookups."""
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
    win = msa["rows"][start[:, None] + np.arange(k)[None, :]]
    win = win[~(win == ord("-")).any(axis=1)]
    return np.ascontiguousarray(win.reshape(-1)), (np.arange(len(win) + 1, dtype=np.uint64) * k)


def reference_arm(args, rank, world):
    """The reference's own CPU implementation (unmodified SeqAlign + KMC API in oracle/_ref, driven by
    oracle/ref_shim.cpp with PloidyFrost's readCov call pattern), all host threads, bounded sample per step."""
    if rank != 0:
        return
    import tempfile
    from oracle.bindings import Checker
    try:
        ref = Checker("ref")
        kind = "reference"
    except Exception:
        ref = Checker("oracle")
        kind = "port"
    cores = os.cpu_count() or 1
    device = None
    try:
        import torch
        if torch.cuda.is_available():
            device = "cuda:0"
    except Exception:
        pass
    with tempfile.TemporaryDirectory(prefix="pfbench_ref_") as d:
        bb, prefix, info = make_workload(args, 0, 1, True, d, device)
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
            "config": workload_config(args, sample.n_bubbles, info),
            "kmc_lookups_per_s": n_lookups / (ms * 1e-3),
            "cpu_baseline": {"value": val, "unit": "bubbles/s", "cores": cores, "kind": kind,
                             "sample": f"{sample.n_bubbles} bubbles ({n_lookups} k-mer lookups + {n_site_kmers} site k-mer lookups) of the same batch per step"},
            "e2e": {"value": val, "unit": "bubbles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, n_bubbles, info):
    return {"workload": f"configs[1]: synthetic tetraploid {args.genome_mbp:g} Mbp (4 haplotypes, 1% SNP, 0.1% indel), "
                        f"60x-equivalent KMC2 db (k=25, p=9, sig 9, 512 bins), -z 8, M/D/G = 2/-1/-3",
            "batch_bubbles": int(n_bubbles), "db_kmers": (info or {}).get("N"),
            "step": "lookup-A (readCov of entrance + branch unitigs) + SequenceAlignment per bubble + lookup-B (site k-mers of the branching bubbles)",
            "l2": "flushed between timed steps (256 MiB memset); KMC index (>2 GB) exceeds L2",
            "kmc_index": ("hash index partitioned by mix(key) % n_gpus; " +
                          ("other partitions mapped through CUDA IPC, buckets loaded over NVLink inside the lookup kernel"
                           if getattr(args, "peer_active", False) else "queries routed by NCCL all-to-all"))
                         if getattr(args, "sharded_db", False) else "replicated on every GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-mbp", type=float, default=100.0)
    ap.add_argument("--batch", type=int, default=262144, help="bubbles per step per GPU")
    ap.add_argument("--ref-sample", type=int, default=65536, help="bubbles per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=131072, help="bubbles of the cpu_baseline leg")
    ap.add_argument("--low", type=int, default=2)
    ap.add_argument("--up", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--region-rank", type=int, default=None,
                    help="diagnostics: take the batch another rank would take (its region of the genome) on this GPU")
    ap.add_argument("--no-peer-lookup", action="store_true",
                    help="with --sharded-db: keep the route / NCCL all-to-all / scatter path instead of mapping the other partitions "
                         "through CUDA IPC and loading their buckets over NVLink")
    ap.add_argument("--sharded-db", action="store_true",
                    help="partition the KMC index across the ranks (bin % world) and route queries with an NCCL all-to-all "
                         "instead of replicating it (BASELINE config 3 variant)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ploidyfrost_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    t_setup = time.perf_counter()
    db_dir = f"/tmp/pfbench_{os.environ.get('MASTER_PORT', 'solo')}_{args.genome_mbp:g}"
    os.makedirs(db_dir, exist_ok=True)
    bb, prefix, info = make_workload(args, rank if args.region_rank is None else args.region_rank, world, rank == 0, db_dir, str(dev))
    if world > 1:
        obj = [info]
        dist.broadcast_object_list(obj, src=0)
        info = obj[0]
    barrier()
    torch.cuda.empty_cache()
    ctx = capi.Context(local_rank)
    if args.sharded_db:
        from ploidyfrost_b200 import sharded
        db = capi.KmcDb(ctx, prefix, part=rank, n_parts=world)
        sh = sharded.ShardedKmcDb(db)
        peer = (not args.no_peer_lookup) and sharded.attach_peers(db, dev)
    else:
        db = capi.KmcDb(ctx, prefix)
        peer = False
    barrier()
    args.peer_active = bool(peer)
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
    do_sites = (not args.sharded_db) or peer                                # lookup phase B needs every k-mer reachable from this GPU
    seq_len = np.diff(bb.seq_off)
    max_len, max_rows = int(seq_len.max()), int(np.diff(bb.bubble_off).max())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # a side stream: the legacy default stream's handle is 0, which the C ABI reads as "use the context's own stream",
    # and torch.cuda.Event only sees the stream it is recorded on -- so kernels and events share this explicit stream
    stream = torch.cuda.Stream(dev)
    sptr = stream.cuda_stream
    assert sptr != 0
    torch.cuda.synchronize()

    def step_device(ev=None):
        if ev:
            ev[0].record(stream)
        if args.sharded_db and not peer:
            sh.lookup(d_lb, d_lo, d_wo, n_win, mode=capi.LOOKUP_CANONICAL, low=args.low, up=args.up, stream=stream)
        else:
            db.lookup_dev(d_lb.data_ptr(), len(lb), d_lo.data_ptr(), d_wo.data_ptr(), n_lseq, n_win, capi.LOOKUP_CANONICAL, args.low,
                          args.up, None, None, d_cov.data_ptr(), sptr)
        if ev:
            ev[1].record(stream)
        ctx.align_dev(d_ab.data_ptr(), len(bb.bases), d_ao.data_ptr(), bb.n_seq, d_bo.data_ptr(), bb.n_bubbles, max_len, max_rows,
                      stream=sptr)
        if ev:
            ev[2].record(stream)
        if do_sites:
            db.site_cov_dev(args.low, args.up, d_skip.data_ptr(), sptr)
        if ev:
            ev[3].record(stream)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 1)):
            step_device()
    torch.cuda.synchronize()
    cells = ctx.last_cells
    retry = ctx.last_retry_count
    heavy_q = ctx.last_heavy_queued

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    with torch.cuda.stream(stream):
        for it in range(args.steps):
            flush.fill_(it & 0xFF)
            step_device(evs[it])
    torch.cuda.synchronize()
    barrier()
    launches = ctx.launches - launches0
    t_lookup = [e[0].elapsed_time(e[1]) for e in evs]
    t_align = [e[1].elapsed_time(e[2]) for e in evs]
    t_site = [e[2].elapsed_time(e[3]) for e in evs]
    t_step = [e[0].elapsed_time(e[3]) for e in evs]
    ms_step = sum(t_step) / len(t_step)

    # ---- e2e: host-pointer C ABI from pinned host buffers ----
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    h_lb, h_lo, h_ab, h_ao, h_bo, h_wo = [pinned(x) for x in (lb, lo, bb.bases, bb.seq_off, bb.bubble_off, wo)]
    keep += [h_lb, h_lo, h_ab, h_ao, h_bo, h_wo]
    e2e_times = []
    e2e_cov_times = []
    cov_pinned = torch.empty(n_lseq * 24, dtype=torch.uint8).pin_memory()
    cov_out = cov_pinned.numpy().view(capi.COV_DTYPE)
    d2h_bytes = 0
    for it in range(2 + args.steps):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.sharded_db and not peer:   # no host-pointer form of the partitioned lookup: copy in, route/exchange/lookup, copy the cov records out
            with torch.cuda.stream(stream):
                e_lb = h_lb[0].to(dev, non_blocking=True)
                e_lo = h_lo[0].to(dev, non_blocking=True).view(torch.int64)
                e_wo = h_wo[0].to(dev, non_blocking=True).view(torch.int64)
                _, _, e_cov = sh.lookup(e_lb, e_lo, e_wo, n_win, mode=capi.LOOKUP_CANONICAL, low=args.low, up=args.up, stream=stream)
                cov_pinned.copy_(e_cov[:n_lseq * 24], non_blocking=True)
            stream.synchronize()
            cov = cov_out
        else:   # lookup-A runs on the handle's own stream beside the alignment (pf_kmc_cov_async ... pf_kmc_wait)
            cov = db.cov_async(h_lb[1], h_lo[1], cov_out, mode=capi.LOOKUP_CANONICAL, low=args.low, up=args.up)
        t1 = time.perf_counter()
        msa = ctx.align(h_ab[1], h_ao[1], h_bo[1], copy=False)   # views of the pinned result arena (C-ABI ownership rule)
        sites = db.site_cov(args.low, args.up, skip_np, copy=False) if do_sites else None   # views, like the alignment result
        if not args.sharded_db or peer:
            db.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= 2:
            e2e_times.append(dt)
            e2e_cov_times.append(t1 - t0)
        d2h_bytes = cov.nbytes + sum(v.nbytes for v in msa.values() if isinstance(v, np.ndarray))
        if sites is not None:
            d2h_bytes += sum(v.nbytes for v in sites.values())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_e2e = 1e3 * sum(e2e_times) / len(e2e_times)
    h2d_bytes = lb.nbytes + lo.nbytes + wo.nbytes + bb.bases.nbytes + bb.seq_off.nbytes + bb.bubble_off.nbytes + (skip_np.nbytes if do_sites else 0)
    site_hist = np.bincount(sites["status"], minlength=5).tolist() if sites is not None else None
    n_sites = int(len(sites["status"])) if sites is not None else 0
    n_ok = int((msa["status"] == 0).sum())

    # ---- reduce over ranks (max time, summed work) ----
    sys.stderr.write(f"[rank {rank}] step {ms_step:.2f} ms (lookup {sum(t_lookup) / len(t_lookup):.2f}, align {sum(t_align) / len(t_align):.2f}, sites {sum(t_site) / len(t_site):.2f}), "
                     f"e2e {ms_e2e:.2f} ms; tiers {ctx.last_tier_counts} heavy-queued {heavy_q}; batch {bb.stats()}\n")
    from ploidyfrost_b200 import shard
    (ms_step, ms_e2e, ms_lookup, ms_align, ms_site), (tot_bubbles, tot_win, tot_cells) = shard.reduce_step(
        [ms_step, ms_e2e, sum(t_lookup) / len(t_lookup), sum(t_align) / len(t_align), sum(t_site) / len(t_site)], [bb.n_bubbles, n_win, cells], device=dev)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        gather_gbs = ctx.bench_random_gather(4 << 30)
        int32_gops = ctx.bench_int32()
        # measured DRAM traffic per launch (ncu, profiles/r01_traffic.json) -- only valid for the workload it was captured on
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        except Exception:
            pass
        default_wl = args.batch == 262144 and args.genome_mbp == 100.0 and not args.sharded_db
        t_lookup = traffic.get("kmc_hash_lookup_kernel", {}).get("bytes") if default_wl and db.index_kind == "hash" else None
        t_align = traffic.get("align_pipeline", {}).get("bytes") if default_wl else None
        lookups_s = (tot_win / world) / (ms_lookup * 1e-3)       # per GPU, for the per-kernel roofline
        cells_s = (tot_cells / world) / (ms_align * 1e-3)
        roof_lookup = {"kernel": "kmc_hash_lookup_kernel" if db.index_kind == "hash" else "kmc_lookup_kernel", "bound": "hbm",
                       "achieved": lookups_s * 64 / 1e9, "peak": hbm_peak,
                       "unit": "GB/s", "frac": lookups_s * 64 / 1e9 / hbm_peak, "traffic": t_lookup, "peak_source": peak_src,
                       "algorithmic_bytes_per_lookup": 64, "lookups_per_launch": tot_win / world, "ms": ms_lookup,
                       "random_sector_gather_gbs": gather_gbs, "index": db.index_kind,
                       "index_bytes": db.device_bytes, "frac_of_random_gather": lookups_s * 64 / 1e9 / gather_gbs}
        roof_align = {"kernel": "alignment pipeline: msa_lane_kernel (branches <= 96) + msa_group_kernel<2/2/4/32> (<= 128 / 192 / 256 / longer), concurrent streams",
                      "bound": "int32", "achieved": cells_s * 18 / 1e9, "peak": int32_gops,
                      "unit": "Gop/s", "frac": cells_s * 18 / 1e9 / int32_gops, "traffic": t_align,
                      "peak_source": "measured in this run (pf_bench_int32: IADD/IMNMX/LOP mix)", "int32_ops_per_cell": 18,
                      "cells_per_launch": tot_cells / world, "ms": ms_align}
        dominant = roof_align if ms_align >= ms_lookup else roof_lookup
        line = {"metric": "superbubble variants/sec", "value": tot_bubbles / (ms_step * 1e-3), "unit": "bubbles/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": workload_config(args, bb.n_bubbles, info),
                "kmc_lookups_per_s": tot_win / (ms_step * 1e-3), "kmc_lookups_per_s_kernel": tot_win / (ms_lookup * 1e-3),
                "dp_cells_per_s_kernel": tot_cells / (ms_align * 1e-3),
                "ms_lookup_kernel": ms_lookup, "ms_align_pipeline": ms_align, "ms_site_cov_kernel": ms_site,
                "site_columns_per_step": n_sites, "site_status_hist(ok,dropped,missing,undefined,skipped)": site_hist,
                "clocks": sampler.summary(),
                "e2e": {"value": tot_bubbles / (ms_e2e * 1e-3), "unit": "bubbles/s", "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e,
                        "ms_pf_kmc_cov_async_enqueue": 1e3 * sum(e2e_cov_times) / len(e2e_cov_times),
                        "calls": "pf_kmc_cov_async | pf_align | pf_site_cov | pf_kmc_wait"},
                "gpu_launches": int(launches),
                "roofline": dominant, "roofline_lookup": roof_lookup, "roofline_align": roof_align,
                "bubbles_ok": n_ok, "tier2_retries": int(retry), "heavy_queued": int(heavy_q), "setup_s": round(t_setup, 1),
                "batch_stats": bb.stats()}
        if not args.no_cpu_baseline and world >= 1:
            line["cpu_baseline"] = cpu_baseline(args, bb, prefix)
        print(json.dumps(line), flush=True)
        for ext in (".kmc_pre", ".kmc_suf"):
            try:
                os.remove(prefix + ext)
            except OSError:
                pass
    db.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, bb, prefix):
    """oracle/_ref (the unmodified reference) on this box's host cores, bounded sample of the same batch, same database."""
    from oracle.bindings import Checker
    try:
        ref = Checker("ref")
        kind = "reference"
    except Exception:
        ref = Checker("oracle")
        kind = "port"
    cores = os.cpu_count() or 1
    sample = bb.slice(0, min(bb.n_bubbles, args.cpu_sample))
    lb, lo = sample.lookup_sequences()
    n_lookups = int(np.maximum(np.diff(lo).astype(np.int64) - K + 1, 0).sum())
    h = ref.kmc_open(prefix)
    t0 = time.perf_counter()
    ref.kmc_cov(h, lb, lo, mode=1, low=args.low, up=args.up, n_threads=cores)
    t1 = time.perf_counter()
    msa = ref.align(sample.bases, sample.seq_off, sample.bubble_off, n_threads=cores)
    t2 = time.perf_counter()
    sb, so = site_kmer_proxy(msa, sample.bubble_type, K)               # untimed: stands in for the reference's string handling
    t3 = time.perf_counter()
    if len(so) > 1:
        ref.kmc_counts(h, sb, so, K, mode=1, n_threads=cores)          # lookup-B
    t4 = time.perf_counter()
    ref.kmc_close(h)
    total = (t2 - t0) + (t4 - t3)
    return {"value": sample.n_bubbles / total, "unit": "bubbles/s", "cores": cores, "kind": kind,
            "sample": f"{sample.n_bubbles} bubbles / {n_lookups} k-mer lookups + {len(so) - 1} site k-mer lookups of the same batch, one pass, same KMC db",
            "kmc_lookups_per_s": n_lookups / (t1 - t0), "align_bubbles_per_s": sample.n_bubbles / (t2 - t1),
            "site_kmer_lookups_per_s": (len(so) - 1) / max(t4 - t3, 1e-9), "seconds": round(total, 2)}


if __name__ == "__main__":
    main()
