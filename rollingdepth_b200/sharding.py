"""Snippet sharding over the GPUs of one box (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no parallelism at all; the units of its hot loops are independent
(each snippet needs only the rgb latents and the shared noise), so every compute stage is block-partitioned over
the ranks and results are exchanged with ONE in-place all-gather per stage into a pre-sized buffer:

  stage                         partition                     exchange
  VAE encode                    frames                        all-gather rgb latents              (6 MB, fast/150 f)
  snippet denoise + decode      snippets of each dilation     all-gather of the decoded snippets  (0.5-1.3 GB over NVLink:
                                                              1-2 ms at the measured 600-700 GB/s)
  co-alignment                  replicated (default): every rank holds every snippet after the gather and runs the
                                single-GPU kernels -- global min, subsample, optimiser, merge -- so the result is
                                bit-identical to a one-GPU run by construction and `snippet_ls` is complete on every rank;
                                or frame-sharded optimiser ("allreduce"): each rank owns a block of frames and the
                                per-(frame, contribution) gradient terms [N, G*w, 2] are combined by an all-reduce per
                                iteration (rd_coalign_begin / pass / exchange / end) -- wins for long videos, where one
                                pass costs more than the all-reduce latency
  refine                        windows of each step          all-gather of the step's predicted latents
  final decode                  frames                        all-gather

Partition: rank r owns the block [r*cap, min(n, (r+1)*cap)) with cap = ceil(n / world).  The largest block is the
same as for a balanced partition (so is the makespan), and because every block but the last ones is full, the
all-gather of per-rank slots of `cap` rows lands every row at its global index -- no padding copies, no list of
buffers, no concatenation.

Works with any backend (`gloo` on CPU for the tests, `nccl` on GPUs).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch


class Shard:
    def __init__(self, rank: int = 0, world: int = 1, group=None):
        self.rank, self.world, self.group = rank, world, group
        self._inplace_ok: Optional[bool] = None

    @classmethod
    def from_dist(cls, group=None) -> "Shard":
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return cls()
        return cls(dist.get_rank(group), dist.get_world_size(group), group)

    # -------------------------------------------------------------- partition
    @staticmethod
    def cap_of(n: int, world: int) -> int:
        return (n + world - 1) // world if n > 0 else 0

    @classmethod
    def block_of(cls, n: int, rank: int, world: int) -> Tuple[int, int]:
        """Contiguous block [lo, hi) of `n` units for `rank`: slots of ceil(n / world) units, the last ones short."""
        cap = cls.cap_of(n, world)
        lo = min(n, rank * cap)
        return lo, min(n, lo + cap)

    def cap(self, n: int) -> int:
        return self.cap_of(n, self.world)

    def block(self, n: int) -> Tuple[int, int]:
        return self.block_of(n, self.rank, self.world)

    def sizes(self, n: int) -> List[int]:
        return [self.block_of(n, r, self.world)[1] - self.block_of(n, r, self.world)[0] for r in range(self.world)]

    # -------------------------------------------------------------- collectives
    def broadcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)
        return t

    def all_reduce_(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
        if self.world > 1:
            import torch.distributed as dist
            ops = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
            dist.all_reduce(t, op=ops[op], group=self.group)
        return t

    def gather_buffer(self, n: int, tail_shape: Sequence[int], dtype, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """(buffer [world * cap(n), *tail], this rank's block view [hi - lo, *tail]) -- fill the view, then `gather_`."""
        rows = max(self.world * self.cap(n), 1) if self.world > 1 else max(n, 0)
        buf = torch.empty((rows, *tail_shape), dtype=dtype, device=device)
        lo, hi = self.block(n)
        return buf, buf[lo:hi]

    def gather_(self, buf: torch.Tensor, n: int) -> torch.Tensor:
        """In-place all-gather of the per-rank slots of `buf` (from `gather_buffer`); returns buf[:n] = all n rows in
        global order on every rank."""
        if self.world > 1 and n > 0:
            import torch.distributed as dist
            cap = self.cap(n)
            mine = buf[self.rank * cap: (self.rank + 1) * cap]
            if dist.get_backend(self.group) != "nccl":
                mine = mine.clone()          # gloo: no aliasing guarantee for in-place all-gather
            dist.all_gather_into_tensor(buf, mine, group=self.group)
        return buf[:n]


def balanced_chunks(n: int, max_chunk: int) -> List[Tuple[int, int]]:
    """[lo, hi) ranges covering n units in ceil(n / max_chunk) chunks whose sizes differ by at most one
    (19 -> 10 + 9 instead of 16 + 3: a tail batch of 3 runs far below the efficiency of a full one)."""
    if n <= 0:
        return []
    k = (n + max_chunk - 1) // max_chunk
    base, rem = divmod(n, k)
    out, lo = [], 0
    for i in range(k):
        hi = lo + base + (1 if i < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


# ---------------------------------------------------------------------------------------------------------------
# Frame-sharded optimiser: per-iteration all-reduce of the per-(frame, contribution) gradient terms
# ---------------------------------------------------------------------------------------------------------------
def coalign_optimize_allreduce(aligner, shard: Shard, pb, x_small: torch.Tensor, state: List[torch.Tensor],
                               n_iter: int) -> torch.Tensor:
    """depth_aligner.py:123-229 with the frames block-partitioned over the ranks: every rank
    holds all subsampled snippets `x_small` (identical on every rank) and evaluates the fused pass for its own
    frames; the only exchange is a sum-all-reduce of the [N, G*w, 2] gradient terms after every pass (each entry
    is written by exactly one rank and is zero elsewhere, so the sum is exact and the parameters are bit-identical
    to a single-GPU run).  Updates `state` in place on every rank; returns the history tensor [n_iter, 3]."""
    from . import _lib as L
    lib = L.lib()
    dev = x_small.device
    hp = aligner._hparams()
    ws_bytes = lib.rd_coalign_workspace_bytes(C.byref(pb), n_iter)
    if ws_bytes < 0:
        L.check(int(ws_bytes), "rd_coalign_workspace_bytes")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = L.stream_ptr(dev)
    f_lo, f_hi = shard.block(pb.N)
    with L.device_guard(dev):
        L.check(lib.rd_coalign_begin(C.byref(pb), *[L.ptr(a) for a in state], n_iter, 0, L.ptr(ws), ws_bytes, st),
                "rd_coalign_begin")
        send, recv, nfl = C.c_void_p(), C.c_void_p(), C.c_int64()
        for it in range(n_iter):
            if f_hi > f_lo:
                L.check(lib.rd_coalign_pass(L.ptr(x_small), C.byref(pb), C.byref(hp), it, 0, f_lo, f_hi, L.ptr(ws), st),
                        "rd_coalign_pass")
                L.count_launches(1)
            L.check(lib.rd_coalign_exchange(C.byref(pb), L.ptr(ws), it, C.byref(send), C.byref(recv), C.byref(nfl)),
                    "rd_coalign_exchange")
            o_s, o_r = send.value - ws.data_ptr(), recv.value - ws.data_ptr()
            t_send = ws[o_s: o_s + 4 * nfl.value].view(torch.float32)
            t_recv = ws[o_r: o_r + 4 * nfl.value].view(torch.float32)
            t_recv.copy_(t_send)
            shard.all_reduce_(t_recv, "sum")
        hist = torch.zeros((max(n_iter, 1), 3), dtype=torch.float32, device=dev)
        out = [torch.zeros_like(a) for a in state]
        if f_hi > f_lo:
            L.check(lib.rd_coalign_end(C.byref(pb), C.byref(hp), n_iter, 0, f_lo, f_hi, *[L.ptr(a) for a in out],
                                       L.ptr(hist), L.ptr(ws), st), "rd_coalign_end")
            L.count_launches(2)
        else:
            hist[:, 1] = float("inf")
            hist[:, 2] = float("-inf")
    # every snippet's final state comes from the rank owning its first frame (zeros elsewhere): sum == assemble
    packed = torch.stack(out)
    shard.all_reduce_(packed, "sum")
    for a, b in zip(state, packed):
        a.copy_(b)
    loss, mn, mx = hist[:, 0].contiguous(), hist[:, 1].contiguous(), hist[:, 2].contiguous()
    shard.all_reduce_(loss, "sum")
    shard.all_reduce_(mn, "min")
    shard.all_reduce_(mx, "max")
    return torch.stack([loss, mn, mx], dim=1)[:n_iter]
