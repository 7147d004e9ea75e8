"""KMC database partitioned across the GPUs of a box (BASELINE config 3).

Each rank opens ONE partition of the database (KMC2: bins with bin % world == rank; KMC1: a prefix range) and
answers the queries that fall into it.  Per batch: route (CUDA) -> all-to-all of 8-byte keys -> lookup at the
owner (CUDA) -> all-to-all of replies -> scatter back into window order + readCov reductions (CUDA).
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the transport; the arithmetic is
in libpfgpu.so (pf_kmc_route_dev / pf_kmc_lookup_keys_dev / pf_kmc_scatter_dev).  Results are identical to the
replicated index.

`exchange()` is transport only and is what the world_size-2 gloo test drives with a CPU stand-in engine.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def exchange(send: torch.Tensor, send_counts, group=None):
    """Variable all-to-all of a 1-D tensor bucketed by destination rank.

    send_counts[r] elements go to rank r (buckets are consecutive in `send`).  Returns (recv, recv_counts):
    what every rank sent to us, in source-rank order.  world_size 1 (or no process group) is the identity."""
    counts = [int(c) for c in send_counts]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return send[:counts[0]], counts[:1]
    world = dist.get_world_size(group)
    assert len(counts) == world
    sc = torch.tensor(counts, dtype=torch.int64, device=send.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send[:sum(counts)].contiguous(), output_split_sizes=recv_counts, input_split_sizes=counts,
                           group=group)
    return recv, recv_counts


def exchange_back(reply: torch.Tensor, recv_counts, send_counts, group=None):
    """The way back: replies for the keys we received (source-rank order) return to their senders; the result is in
    our original send order."""
    rc, sc = [int(c) for c in recv_counts], [int(c) for c in send_counts]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return reply
    out = torch.empty(sum(sc), dtype=reply.dtype, device=reply.device)
    dist.all_to_all_single(out, reply.contiguous(), output_split_sizes=sc, input_split_sizes=rc, group=group)
    return out


def attach_peers(db, device, group=None) -> bool:
    """Peer-memory form: every rank exports its slice of the hash index (CUDA IPC handle), the blobs are all-gathered and the
    other slices are mapped (pf_kmc_attach_peers).  After that db.lookup_dev / db.cov answer every query in ONE kernel by
    loading the owner's bucket over NVLink.  Returns False (and leaves the exchange path in place) when the partition is not a
    hash slice or the mapping is refused."""
    from . import capi
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or db.index_kind != "hash":
        return False
    world = dist.get_world_size(group)
    ok = 1
    try:
        mine = torch.from_numpy(db.export_ipc()).to(device)
    except capi.PfError:
        mine, ok = torch.zeros(128, dtype=torch.uint8, device=device), 0
    blobs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(blobs, mine, group=group)
    if ok:
        try:
            db.attach_peers(torch.cat(blobs).cpu().numpy())
        except capi.PfError:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # all ranks or none
    return bool(flag.item())


class ShardedKmcDb:
    """One rank's view of a partitioned KMC database.  `db` is a capi.KmcDb opened with (part=rank, n_parts=world)."""

    def __init__(self, db, group=None):
        self.db = db
        self.group = group
        self.k = db.k
        self.last_sent = 0
        self.last_received = 0
        self._stream = None   # the stream lookup() runs on when the caller gives none (never torch's default stream)

    def lookup(self, d_bases: torch.Tensor, d_seq_off: torch.Tensor, d_win_off: torch.Tensor, n_windows: int, mode=0, low=0,
               up=0xFFFFFFFF, want_cov=True, stream=None):
        """Device tensors in (uint8 bases -- may be padded past the last sequence --, int64 offsets), device tensors
        out: (counts int32-as-u32, found uint8, cov bytes)."""
        dev = d_bases.device
        # The C calls and the NCCL exchanges must share ONE explicit stream: handle 0 (torch's legacy default stream) means "the
        # context's own non-blocking stream" to the C ABI, which NCCL's collectives (ordered on torch's current stream) would not
        # order against.  So the body always runs under a non-default torch stream; when the caller did not supply one, ours is
        # ordered after the caller's current stream (the inputs) and the caller's stream after ours (the outputs).
        own = stream is None or stream.cuda_stream == 0
        if own:
            if self._stream is None:
                self._stream = torch.cuda.Stream(dev)
            stream = self._stream
            stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            out = self._lookup_on(stream, d_bases, d_seq_off, d_win_off, n_windows, mode, low, up, want_cov)
        if own:
            cur = torch.cuda.current_stream(dev)
            cur.wait_stream(stream)
            for t in (d_bases, d_seq_off, d_win_off):   # allocated on the caller's stream, read on ours
                t.record_stream(stream)
            for t in out:                                # allocated on ours, read on the caller's
                if t is not None:
                    t.record_stream(cur)
        return out

    def _lookup_on(self, stream, d_bases, d_seq_off, d_win_off, n_windows, mode, low, up, want_cov):
        dev = d_bases.device
        n_seq = d_seq_off.numel() - 1
        sptr = stream.cuda_stream
        assert sptr != 0
        send_keys = torch.empty(max(n_windows, 1), dtype=torch.int64, device=dev)
        send_idx = torch.empty(max(n_windows, 1), dtype=torch.int32, device=dev)
        off = self.db.route_dev(d_bases.data_ptr(), d_bases.numel(), d_seq_off.data_ptr(), d_win_off.data_ptr(), n_seq, n_windows, mode,
                                send_keys.data_ptr(), send_idx.data_ptr(), sptr)
        send_counts = np.diff(off).astype(np.int64)
        n_sent = int(off[-1])
        recv_keys, recv_counts = exchange(send_keys, send_counts, self.group)
        n_recv = int(sum(recv_counts))
        r_counts = torch.empty(max(n_recv, 1), dtype=torch.int32, device=dev)
        r_found = torch.empty(max(n_recv, 1), dtype=torch.uint8, device=dev)
        self.db.lookup_keys_dev(recv_keys.data_ptr(), n_recv, r_counts.data_ptr(), r_found.data_ptr(), sptr)
        b_counts = exchange_back(r_counts[:n_recv], recv_counts, send_counts, self.group)
        b_found = exchange_back(r_found[:n_recv], recv_counts, send_counts, self.group)
        counts = torch.empty(max(n_windows, 1), dtype=torch.int32, device=dev)
        found = torch.empty(max(n_windows, 1), dtype=torch.uint8, device=dev)
        cov = torch.empty(max(n_seq, 1) * 24, dtype=torch.uint8, device=dev) if want_cov else None
        self.db.scatter_dev(send_idx.data_ptr(), n_sent, b_counts.data_ptr(), b_found.data_ptr(), d_win_off.data_ptr(), n_seq, n_windows,
                            low, up, counts.data_ptr(), found.data_ptr(), cov.data_ptr() if want_cov else None, sptr)
        self.last_sent, self.last_received = n_sent, n_recv
        return counts[:n_windows], found[:n_windows], cov
