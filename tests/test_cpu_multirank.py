"""world_size-2 gloo test of the N>1 path: bubbles shard by rank with no data-path collective; shards
concatenate in rank order to exactly the unsharded result; timings reduce as max, work as sum."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.bindings import Checker, flatten_bubbles
from ploidyfrost_b200 import shard
from tests import gen
from tests.util import MSA_KEYS


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bubbles = gen.random_bubbles(77, 501)          # odd count: ranges differ in size
        bases, off, boff = flatten_bubbles(bubbles)
        lb, lo, lbo = shard.shard_batch(bases, off, boff, rank, world)
        orc = Checker("oracle")                        # stand-in engine on CPU; the GPU engine has the same call shape
        m = orc.align(lb, lo, lbo)
        sizes = shard.gather_sizes([m["n_bubbles"], len(m["rows"]), len(m["var_col"])])
        times, work = shard.reduce_step([10.0 + rank, 5.0 - rank], [m["n_bubbles"], 1])
        np.savez(os.path.join(outdir, f"r{rank}.npz"), sizes=sizes, times=times, work=work, **{k: m[k] for k in MSA_KEYS})
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_bubble_ranges_cover_exactly():
    for n in (0, 1, 7, 64, 501):
        for world in (1, 2, 3, 8):
            r = [shard.bubble_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        shard.bubble_range(5, 2, 2)


def test_two_ranks_gloo_shards_equal_unsharded(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [dict(np.load(os.path.join(str(tmp_path), f"r{r}.npz"))) for r in range(world)]
    for p in parts:
        p["n_bubbles"] = len(p["n_rows"])
    bubbles = gen.random_bubbles(77, 501)
    whole = oracle.align(*flatten_bubbles(bubbles))
    merged = shard.concat_msa(parts)
    for k in MSA_KEYS:
        assert np.array_equal(np.asarray(whole[k]), merged[k]), k
    # every rank saw the same size table, in rank order, and the reduced stats
    assert np.array_equal(parts[0]["sizes"], parts[1]["sizes"])
    assert list(parts[0]["sizes"][:, 0]) == [251, 250]
    assert int(parts[0]["sizes"][:, 1].sum()) == len(whole["rows"])
    for p in parts:
        assert list(p["times"]) == [11.0, 5.0] and list(p["work"]) == [501.0, 2.0]


def _exchange_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ploidyfrost_b200 import sharded
        rng = np.random.default_rng(100 + rank)
        keys = rng.integers(0, 1 << 50, 1000 + 37 * rank, dtype=np.int64)
        owner = (keys % 7) % world                                # stand-in for "bin % world"
        order = np.argsort(owner, kind="stable")
        send = torch.from_numpy(keys[order])
        counts = np.bincount(owner, minlength=world)
        recv, rcounts = sharded.exchange(send, counts)
        assert all(int(x) % 7 % world == rank for x in recv.tolist())       # we only receive keys we own
        reply = recv * 3 + rank                                   # "lookup at the owner"
        back = sharded.exchange_back(reply, rcounts, counts)
        got = np.empty(len(keys), dtype=np.int64)
        got[order] = back.numpy()
        assert np.array_equal(got, keys * 3 + owner)              # answers are back in the original window order
        total = torch.tensor([len(recv)])
        dist.all_reduce(total)
        np.save(os.path.join(outdir, f"x{rank}.npy"), np.array([len(keys), int(total)]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_key_exchange_roundtrip(tmp_path):
    """Transport of the partitioned-database path (sharded.exchange / exchange_back) over gloo, world_size 2."""
    world = 2
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(os.path.join(str(tmp_path), f"x{r}.npy")) for r in range(world))
    assert a[1] == b[1] == a[0] + b[0]      # every key was delivered exactly once
