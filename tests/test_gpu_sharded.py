"""-m gpu: the partitioned KMC index (pf_kmc_open_part / route / lookup_keys / scatter) against the unpartitioned
lookup on the same GPU -- all partitions are opened side by side and the all-to-all is played by slicing, so the
kernels and the partition arithmetic are checked without needing several GPUs.  The NCCL transport itself is
covered by test_two_gpu_sharded_lookup_nccl (needs >= 2 GPUs) and by the gloo test in test_cpu_multirank.py."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle.bindings import flatten_seqs
from tests import gen

pytestmark = pytest.mark.gpu


def _dev_batch(seqs, k, dev):
    from ploidyfrost_b200 import capi
    bases, off = flatten_seqs(seqs)
    wo = capi.window_offsets(off, k)
    pad = (-len(bases)) % 16
    d_b = torch.from_numpy(np.concatenate([bases, np.zeros(pad, np.uint8)])).to(dev)
    return bases, off, wo, d_b, torch.from_numpy(off.astype(np.int64)).to(dev), torch.from_numpy(wo.astype(np.int64)).to(dev)


@pytest.mark.parametrize("index", ["auto", "verbatim"])
@pytest.mark.parametrize("ver,p,n_parts", [(0x200, 5, 4), (0x200, 9, 3), (0, 5, 4), (0, 1, 8), (0, 9, 5), (0x200, 5, 1)])
def test_partitioned_lookup_equals_replicated(gpu_ctx, tmp_path, ver, p, n_parts, index):
    from ploidyfrost_b200 import capi
    k = 25
    dev = torch.device("cuda", 0)
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=31 + p, k=k, version=ver, p=p, genome_len=60000, n_bins=37)
    rng = np.random.default_rng(p)
    seqs = gen.query_sequences(rng, g, 3000, k=k) + ["", "ACGT", "N" * 30, g[:2000]]
    bases, off, wo, d_b, d_o, d_w = _dev_batch(seqs, k, dev)
    nw, ns = int(wo[-1]), len(off) - 1
    full = capi.KmcDb(gpu_ctx, prefix)
    parts = [capi.KmcDb(gpu_ctx, prefix, part=r, n_parts=n_parts, index=index) for r in range(n_parts)]
    try:
        assert all(d.index_kind == ("hash" if index == "auto" else "verbatim") for d in parts)
        assert sum(d.local_kmers for d in parts) == full.info["total_kmers"]
        if n_parts > 1:
            assert max(d.local_kmers for d in parts) < full.info["total_kmers"]
            with pytest.raises(capi.PfError):
                parts[0].counts(bases, off)          # a partition cannot answer the replicated call
        for mode in (0, 1, 2):
            exp_c, exp_f = full.counts(bases, off, mode=mode)
            exp_cov = full.cov(bases, off, mode=mode, low=1, up=3)
            keys = torch.empty(nw, dtype=torch.int64, device=dev)
            idx = torch.empty(nw, dtype=torch.int32, device=dev)
            soff = parts[0].route_dev(d_b.data_ptr(), d_b.numel(), d_o.data_ptr(), d_w.data_ptr(), ns, nw, mode, keys.data_ptr(), idx.data_ptr())
            n_sent = int(soff[-1])
            assert n_sent <= nw and (np.diff(soff.astype(np.int64)) >= 0).all()
            rc = torch.zeros(max(n_sent, 1), dtype=torch.int32, device=dev)
            rf = torch.zeros(max(n_sent, 1), dtype=torch.uint8, device=dev)
            for r in range(n_parts):                  # "all-to-all": partition r answers its bucket
                a, b = int(soff[r]), int(soff[r + 1])
                if b > a:
                    parts[r].lookup_keys_dev(keys[a:b].data_ptr(), b - a, rc[a:b].data_ptr(), rf[a:b].data_ptr())
            cnt = torch.empty(nw, dtype=torch.int32, device=dev)
            fnd = torch.empty(nw, dtype=torch.uint8, device=dev)
            cov = torch.empty(ns * 24, dtype=torch.uint8, device=dev)
            parts[0].scatter_dev(idx.data_ptr(), n_sent, rc.data_ptr(), rf.data_ptr(), d_w.data_ptr(), ns, nw, 1, 3, cnt.data_ptr(),
                                 fnd.data_ptr(), cov.data_ptr())
            gpu_ctx.sync()
            torch.cuda.synchronize()
            assert np.array_equal(cnt.cpu().numpy().view(np.uint32), exp_c), mode
            assert np.array_equal(fnd.cpu().numpy(), exp_f), mode
            assert np.array_equal(cov.cpu().numpy().view(capi.COV_DTYPE), exp_cov), mode
            # a key sent to the wrong partition is reported absent, never answered from someone else's records
            if n_parts > 1 and n_sent:
                a, b = int(soff[0]), int(soff[1])
                if b > a:
                    parts[1].lookup_keys_dev(keys[a:b].data_ptr(), b - a, rc[a:b].data_ptr(), rf[a:b].data_ptr())
                    gpu_ctx.sync()
                    assert int(rf[a:b].sum()) == 0
    finally:
        full.close()
        for d in parts:
            d.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, prefix, seqs, k, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from ploidyfrost_b200 import capi, sharded
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ctx = capi.Context(rank)
        db = capi.KmcDb(ctx, prefix, part=rank, n_parts=world)
        sh = sharded.ShardedKmcDb(db)
        mine = seqs[rank::world]                       # every rank queries its own sequences
        bases, off, wo, d_b, d_o, d_w = _dev_batch(mine, k, dev)
        cnt, fnd, cov = sh.lookup(d_b, d_o, d_w, int(wo[-1]), mode=0, low=1, up=3)
        torch.cuda.synchronize()
        # peer-memory form: map the other partition through CUDA IPC, then the ordinary host-pointer calls answer everything
        peer = sharded.attach_peers(db, dev)
        p_cnt = p_fnd = np.zeros(0)
        p_cov = np.zeros(0, np.uint8)
        if peer:
            for mode in (0, 1, 2):
                pc, pf = db.counts(bases, off, mode=mode)
                if mode == 0:
                    p_cnt, p_fnd = pc, pf
                    p_cov = db.cov(bases, off, mode=0, low=1, up=3).view(np.uint8)
                np.savez(os.path.join(outdir, f"peer{rank}_m{mode}.npz"), cnt=pc, fnd=pf)
            # lookup phase B through the peers: the end-to-end fixture of the reference, index split over the two GPUs
            from tests import e2e_rows
            meta, eb, _, _ = e2e_rows.load_fixture()
            edb = capi.KmcDb(ctx, os.path.join(e2e_rows.E2E, "db"), part=rank, n_parts=world)
            assert sharded.attach_peers(edb, dev)
            hook, _ = e2e_rows.device_site_cov_hook(edb, meta, eb)
            n_rows, _ = e2e_rows.check_against_reference(lambda *f: ctx.align(*f),
                                                         lambda b, o: edb.cov(b, o, mode=capi.LOOKUP_FWD_THEN_RC, low=2, up=1000), None, hook)
            assert n_rows == 493
            dist.barrier()
            edb.close()
        np.savez(os.path.join(outdir, f"r{rank}.npz"), cnt=cnt.cpu().numpy().view(np.uint32), fnd=fnd.cpu().numpy(),
                 cov=cov.cpu().numpy(), sent=sh.last_sent, recv=sh.last_received, peer=int(peer), p_cnt=p_cnt, p_fnd=p_fnd, p_cov=p_cov)
        dist.barrier()
        db.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_lookup_nccl(gpu_ctx, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from ploidyfrost_b200 import capi
    k, world = 25, 2
    prefix, g, u, c = gen.make_genome_db(tmp_path, seed=5, k=k, version=0x200, p=5, genome_len=80000, n_bins=64)
    rng = np.random.default_rng(1)
    seqs = gen.query_sequences(rng, g, 2000, k=k)
    mp.spawn(_nccl_worker, args=(world, _free_port(), prefix, seqs, k, str(tmp_path)), nprocs=world, join=True)
    full = capi.KmcDb(gpu_ctx, prefix)
    try:
        for r in range(world):
            z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
            bases, off = flatten_seqs(seqs[r::world])
            ec, ef = full.counts(bases, off, mode=0)
            assert np.array_equal(z["cnt"], ec) and np.array_equal(z["fnd"], ef)
            assert np.array_equal(z["cov"].view(capi.COV_DTYPE), full.cov(bases, off, mode=0, low=1, up=3))
            assert int(z["sent"]) > 0 and int(z["recv"]) > 0
            assert int(z["peer"]) == 1, "CUDA IPC mapping of the other partition was refused"
            assert np.array_equal(z["p_cnt"], ec) and np.array_equal(z["p_fnd"], ef)
            assert np.array_equal(z["p_cov"].view(capi.COV_DTYPE), full.cov(bases, off, mode=0, low=1, up=3))
            for mode in (1, 2):
                zz = np.load(os.path.join(str(tmp_path), f"peer{r}_m{mode}.npz"))
                mc, mf = full.counts(bases, off, mode=mode)
                assert np.array_equal(zz["cnt"], mc) and np.array_equal(zz["fnd"], mf), mode
    finally:
        full.close()
