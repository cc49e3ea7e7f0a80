"""Hash-range sharded `compute` across the GPUs of one box (the reference is single-process).

    rank r:  partition   the k-mers of its slice of window positions, grouped by owner rank        (kc_shard_partition)
             exchange    counts, then (scrambled k-mer, global position) items: all-to-all over NVLink   (NCCL)
             resolve     every occurrence of its hash range: first occurrence per kept k-mer -> flag bits  (kc_shard_resolve)
             reduce      the disjoint flag bit arrays onto rank 0 (SUM == OR)                            (NCCL)
    rank 0:  runs -> overlap levels -> superstring (the greedy merge is sequential)                       (kc_compute_from_flags)

The orchestration below is backend-agnostic: `ops` does the per-rank halves and `comm` the collectives.  The product
pairing is GpuOps (libkcgpu through ctypes, torch CUDA tensors as buffers) + TorchComm over NCCL; tests/ pair the same
orchestration with a CPU stand-in for the kernels and TorchComm over gloo to check the host-side logic with
world_size 2 on a machine without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

N_DIGITS = 256  # level-0 groups of the k-mer set construction (top 8 bits of the scrambled word)


def plan_slices(n_bytes: int, world: int, granule: int):
    """Window-END position slices [begin, end) per rank: multiples of `granule`, the last one ends at n_bytes."""
    tiles = -(-n_bytes // granule)
    out = []
    for r in range(world):
        b = min(n_bytes, (tiles * r // world) * granule)
        e = n_bytes if r == world - 1 else min(n_bytes, (tiles * (r + 1) // world) * granule)
        out.append((b, e))
    return out


def owner_of_digit(digit: int, world: int) -> int:
    return digit * world // N_DIGITS


def owner_counts(digit_counts, world: int):
    """Items per owner rank from the 256 digit counts (the owners' digit ranges are contiguous and ascending)."""
    owners = np.arange(N_DIGITS) * world // N_DIGITS
    return np.bincount(owners, weights=np.asarray(digit_counts, dtype=np.float64), minlength=world).astype(np.int64)


@dataclass
class ShardedResult:
    result: object          # rank 0: what ops.finish returned; other ranks: None
    n_kept: int             # distinct k-mers kept (all ranks)
    n_occurrences: int      # k-mer windows seen (all ranks)
    items_sent: int         # items this rank sent to other ranks
    items_received: int     # items this rank resolved


def sharded_compute(ops, comm, n_bytes: int, *, k: int, complements: bool = True, min_frequency: int = 1) -> ShardedResult:
    """One pass of the sharded path.  `ops` already holds the framed sequence of n_bytes bytes."""
    world, rank = comm.world, comm.rank
    b, e = plan_slices(n_bytes, world, ops.granule(k))[rank]
    digit_counts, n_items = ops.partition(b, e, k=k, complements=complements)           # items grouped by owner
    send = owner_counts(digit_counts, world)
    assert int(send.sum()) == n_items
    recv = comm.exchange_counts(send)                                                    # recv[s] = items rank s sends here
    keys, pos = ops.exchange_items(comm, send, recv)                                     # all-to-all (two tensors)
    n_recv = int(recv.sum())
    kept = ops.resolve(keys, pos, n_recv, k=k, complements=complements, min_frequency=min_frequency)
    ops.reduce_flags(comm)                                                               # disjoint bits: SUM == OR, onto rank 0
    tot = comm.sum_scalars([kept, n_items])
    res = ops.finish(int(tot[0]), k=k, complements=complements) if rank == 0 else None
    return ShardedResult(res, int(tot[0]), int(tot[1]), int(n_items - send[rank]), n_recv)


def sharded_compute_p2p(ops, comm, n_bytes: int, *, k: int, complements: bool = True, min_frequency: int = 1,
                        slice_output: bool = False) -> ShardedResult:
    """The product multi-GPU pass: the level-0 scatter stores every item straight into its owner's buffer over NVLink
    (peer pointers), so there is no item all-to-all at all — only the 256 digit counts travel through a collective.
    slice_output: the flags are ALL-reduced, every rank runs the (deterministic, sequential) greedy stage and emits slice
    `rank` of the superstring; result.slice_begin / slice_len say which bytes result.ms_ptr holds on this rank."""
    world, rank = comm.world, comm.rank
    b, e = plan_slices(n_bytes, world, ops.granule(k))[rank]
    counts = ops.p2p_hist(b, e, k=k, complements=complements)                            # 256 digit counts of the slice
    all_counts = comm.all_gather_counts(counts)                                          # [world, 256]
    ops.p2p_scatter(b, e, all_counts, k=k, complements=complements)                      # partition pass == all-to-all
    comm.barrier()                                                                       # every peer's stores have landed
    kept, owned = ops.p2p_resolve(all_counts, k=k, complements=complements, min_frequency=min_frequency)
    ops.reduce_flags(comm, all_ranks=slice_output)
    tot = comm.sum_scalars([kept, int(counts.sum())])                                    # also fences the next pass's stores
    if slice_output:
        res = ops.finish(int(tot[0]), k=k, complements=complements, slice=(rank, world))
    else:
        res = ops.finish(int(tot[0]), k=k, complements=complements) if rank == 0 else None
    mine = int(all_counts[rank][[g for g in range(N_DIGITS) if owner_of_digit(g, world) == rank]].sum())
    return ShardedResult(res, int(tot[0]), int(tot[1]), int(counts.sum()) - mine, int(owned))


class TorchComm:
    """torch.distributed plumbing (NCCL on the GPU box, gloo in the CPU tests)."""

    def __init__(self, device):
        import torch.distributed as dist
        self.dist = dist
        self.device = device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0

    def exchange_counts(self, send):
        import torch
        s = torch.as_tensor(np.asarray(send, dtype=np.int64), device=self.device)
        r = torch.empty_like(s)
        if self.world > 1:
            self.dist.all_to_all_single(r, s)
        else:
            r.copy_(s)
        return r.cpu().numpy()

    def all_gather_counts(self, counts):
        import torch
        c = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=self.device)
        out = torch.empty((self.world, c.numel()), dtype=torch.int64, device=self.device)
        if self.world > 1:
            self.dist.all_gather_into_tensor(out, c)
        else:
            out[0].copy_(c)
        return out.cpu().numpy()

    def all_gather_bytes(self, b: np.ndarray) -> np.ndarray:
        import torch
        t = torch.as_tensor(np.asarray(b, dtype=np.uint8), device=self.device)
        out = torch.empty((self.world, t.numel()), dtype=torch.uint8, device=self.device)
        if self.world > 1:
            self.dist.all_gather_into_tensor(out, t)
        else:
            out[0].copy_(t)
        return out.cpu().numpy()

    def all_to_all(self, out, inp, recv_counts, send_counts, width: int = 1):
        """Variable-size all-to-all of rows of `width` elements."""
        if self.world > 1:
            self.dist.all_to_all_single(out, inp, [int(c) * width for c in recv_counts], [int(c) * width for c in send_counts])
        else:
            out.copy_(inp[:out.numel()])

    def reduce_sum(self, t, dst: int = 0):
        if self.world > 1:
            self.dist.reduce(t, dst=dst, op=self.dist.ReduceOp.SUM)

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)

    def sum_scalars(self, values):
        import torch
        t = torch.as_tensor(np.asarray(values, dtype=np.int64), device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()


class GpuOps:
    """The per-rank halves on the GPU: libkcgpu entry points over torch CUDA buffers (device memory plumbing only)."""

    def __init__(self, ctx, seq_dev):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.seq = seq_dev                      # uint8 CUDA tensor: the whole framed sequence
        self.n_bytes = seq_dev.numel()
        self.flags = torch.zeros((self.n_bytes + 31) // 32 + 1, dtype=torch.int32, device=seq_dev.device)
        self._send_k = self._send_p = None
        self._limbs = 1
        # torch work (copies, memsets, NCCL collectives) is ordered on torch's current stream; the library's kernels are
        # ordered with it only when the context runs on that same stream.  Otherwise every hand-over needs a host sync.
        self._shared_stream = ctx.stream_handle is not None and ctx.stream_handle == torch.cuda.current_stream().cuda_stream

    def _order(self):
        """Make torch-side work visible to the library's stream (no-op when both use one stream)."""
        if not self._shared_stream:
            self.torch.cuda.current_stream().synchronize()

    def granule(self, k):
        return int(self.ctx._lib.kc_shard_granule(k))

    def partition(self, b, e, *, k, complements):
        torch = self.torch
        from .api import limbs_for_k
        self._limbs = limbs_for_k(k)
        self._order()
        cap = max(e - b, 1)
        if self._send_k is None or self._send_k.numel() < cap * self._limbs:
            self._send_k = torch.empty(cap * self._limbs, dtype=torch.int64, device=self.seq.device)
            self._send_p = torch.empty(cap, dtype=torch.int32, device=self.seq.device)
        counts, n = self.ctx.shard_partition(self.seq.data_ptr(), self.n_bytes, b, e, self._send_k.data_ptr(), self._send_p.data_ptr(),
                                             k=k, complements=complements)
        self._n_send = n
        return counts, n

    def exchange_items(self, comm, send, recv):
        torch = self.torch
        n_recv = int(recv.sum())
        L = self._limbs
        keys = torch.empty(max(n_recv, 1) * L, dtype=torch.int64, device=self.seq.device)
        pos = torch.empty(max(n_recv, 1), dtype=torch.int32, device=self.seq.device)
        comm.all_to_all(keys[:n_recv * L], self._send_k[:self._n_send * L], recv, send, L)
        comm.all_to_all(pos[:n_recv], self._send_p[:self._n_send], recv, send, 1)
        return keys, pos

    def resolve(self, keys, pos, n, *, k, complements, min_frequency):
        self.flags.zero_()
        self._order()
        return self.ctx.shard_resolve(keys.data_ptr(), pos.data_ptr(), n, self.flags.data_ptr(), k=k, complements=complements,
                                      min_frequency=min_frequency)

    def reduce_flags(self, comm, all_ranks: bool = False):
        if all_ranks:
            comm.all_reduce_sum(self.flags)
        else:
            comm.reduce_sum(self.flags, 0)

    # ---- fused partition + exchange (peer memory) --------------------------------------------------------------
    def setup_p2p(self, comm, k: int, slack: float = 1.25):
        """Allocate this rank's receive buffers, exchange the CUDA IPC handles, map every peer's buffers."""
        cap = int(self.n_bytes / comm.world * slack) + (1 << 20)
        handles = self.ctx.p2p_alloc(k, cap)
        allh = comm.all_gather_bytes(handles)
        self.ctx.p2p_open(comm.world, comm.rank, allh.reshape(-1))
        comm.barrier()

    def p2p_hist(self, b, e, *, k, complements):
        self._order()
        return self.ctx.p2p_hist(self.seq.data_ptr(), self.n_bytes, b, e, k=k, complements=complements)

    def p2p_scatter(self, b, e, all_counts, *, k, complements):
        self.ctx.p2p_scatter(self.seq.data_ptr(), self.n_bytes, b, e, all_counts, k=k, complements=complements)

    def p2p_resolve(self, all_counts, *, k, complements, min_frequency):
        self.flags.zero_()
        self._order()
        return self.ctx.p2p_resolve(all_counts, self.flags.data_ptr(), k=k, complements=complements, min_frequency=min_frequency)

    def finish(self, n_kept, *, k, complements, slice=None):
        self._order()
        return self.ctx.compute_from_flags(self.seq.data_ptr(), self.n_bytes, self.flags.data_ptr(), n_kept, k=k, complements=complements,
                                           slice=slice)
