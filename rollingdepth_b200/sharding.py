"""Snippet sharding over the GPUs of one box (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no parallelism at all; the units of its hot loops are independent
(each snippet needs only the rgb latents and the shared noise), so every stage is block-partitioned over the
ranks and only small tensors cross NVLink:

  stage                         partition                     exchange
  VAE encode                    frames                        all-gather rgb latents            (~6 MB, fast/150f)
  snippet denoise + decode      snippets of each dilation     none
  co-alignment                  snippets stay where decoded   all-reduce(min) 1 scalar; all-gather of the SUBSAMPLED
                                                              predictions (6-77 MB); optimiser replicated (bit-identical
                                                              on every rank) or frame-sharded with a per-iteration
                                                              all-reduce of the per-(frame, contribution) gradient terms
  merge                         own snippets -> per-frame     all-reduce of fp32 (sum, count) partials [N, H, W]
                                partial sums
  refine                        windows of each step          all-gather of the step's predicted latents
  final decode                  frames                        all-gather

Works with any backend (`gloo` on CPU for the tests, `nccl` on GPUs).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


class Shard:
    def __init__(self, rank: int = 0, world: int = 1, group=None):
        self.rank, self.world, self.group = rank, world, group

    @classmethod
    def from_dist(cls, group=None) -> "Shard":
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return cls()
        return cls(dist.get_rank(group), dist.get_world_size(group), group)

    # -------------------------------------------------------------- partition
    @staticmethod
    def block_of(n: int, rank: int, world: int) -> Tuple[int, int]:
        """Contiguous balanced block [lo, hi) of `n` units for `rank` (sizes differ by at most one)."""
        base, rem = divmod(n, world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)

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

    def all_gather_blocks(self, local: torch.Tensor, total: int) -> torch.Tensor:
        """Concatenate per-rank blocks (block sizes from `block_of(total, r)`) along dim 0 on every rank.
        `local` is either this rank's block or a full-size buffer whose own block is filled in."""
        if self.world == 1:
            return local
        import torch.distributed as dist
        sizes = self.sizes(total)
        lo, hi = self.block(total)
        if local.shape[0] == total:
            mine = local[lo:hi]
        else:
            assert local.shape[0] == hi - lo, (local.shape, lo, hi)
            mine = local
        mx = max(sizes)
        pad = torch.zeros((mx, *mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        pad[: mine.shape[0]] = mine
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def sharded_coalign(aligner, shard: Shard, local_snips: List[torch.Tensor], dilations: List[int], seq_len: int,
                    window: int):
    """Multi-GPU co-alignment: `local_snips[g]` holds this rank's block of the snippets of dilation g
    (block_of(n_g, rank)).  Returns (merged [N,1,H,W], scales, translations, history, local_snips)."""
    import ctypes as C

    from . import _lib as L
    dev = aligner.device
    lib = L.lib()
    dt = local_snips[0].dtype
    local = [s.to(device=dev, dtype=dt).contiguous() for s in local_snips]
    H, W = int(local[0].shape[-2]), int(local[0].shape[-1])

    # global problem (all snippets) and this rank's sub-problem view
    pb = L.CoalignProblem()
    pb.G, pb.w, pb.N, pb.H, pb.W = len(dilations), window, seq_len, H, W
    pb.border, pb.factor = int(aligner.border), int(aligner.factor)
    for g, d in enumerate(dilations):
        pb.d[g] = int(d)
    pb.dtype = L.dtype_code(dt)
    L.check(lib.rd_coalign_problem_init(C.byref(pb)), "rd_coalign_problem_init")
    blocks = [shard.block(pb.n[g]) for g in range(pb.G)]
    for g, s in enumerate(local):
        assert s.shape[0] == blocks[g][1] - blocks[g][0], (g, s.shape, blocks[g])

    st = L.stream_ptr(dev)
    scratch = torch.empty(lib.rd_reduce_scratch_bytes(), dtype=torch.uint8, device=dev)

    # (1) global minimum: local streaming reduction + all-reduce(min) of one scalar
    dev_min = torch.full((1,), float("inf"), dtype=torch.float32, device=dev)
    lp = _local_problem(pb, blocks)
    nonempty = [g for g in range(pb.G) if local[g].shape[0] > 0]
    if nonempty:
        sub, ptrs = _sub_problem(lp, local, nonempty)
        L.check(lib.rd_global_min(ptrs, C.byref(sub), L.ptr(dev_min), L.ptr(scratch), st), "rd_global_min")
    shard.all_reduce_(dev_min, "min")

    # (2) subsample the local snippets, all-gather the small tensors into the global row order
    x_small = torch.zeros((pb.n_total, pb.w, pb.P_pad), dtype=dt, device=dev)
    if nonempty:
        xs_local = torch.empty((sub.n_total, pb.w, pb.P_pad), dtype=dt, device=dev)
        L.check(lib.rd_coalign_prep(ptrs, C.byref(sub), L.ptr(dev_min), L.ptr(xs_local), st), "rd_coalign_prep")
        off = 0
        for g in nonempty:
            n = local[g].shape[0]
            x_small[pb.row0[g] + blocks[g][0]: pb.row0[g] + blocks[g][1]] = xs_local[off:off + n]
            off += n
    shard.all_reduce_(x_small, "sum")    # rows are disjoint and zero elsewhere: x + 0 is exact, so sum == gather

    # (3) optimiser replicated on every rank: identical inputs and a deterministic kernel => identical (s, t)
    state = aligner.new_state(pb)
    hist = aligner.optimize_state(pb, x_small, state, 0, int(aligner.num_iterations))

    # (4) merge: fp32 partial (sum, count) of the local snippets, all-reduce, then round once like the reference
    merged = _sharded_merge(aligner, shard, pb, blocks, local, dev_min, state[0], state[1])
    scales = [state[0][pb.row0[g]: pb.row0[g] + pb.n[g]].reshape(-1, 1, 1) for g in range(pb.G)]
    translations = [state[1][pb.row0[g]: pb.row0[g] + pb.n[g]].reshape(-1, 1, 1) for g in range(pb.G)]
    history = [tuple(r) for r in hist.cpu().tolist()]
    return merged, scales, translations, history, local


def _local_problem(pb, blocks):
    return dict(pb=pb, blocks=blocks)


def _sub_problem(lp, local, groups):
    """A problem struct describing only this rank's snippets (for min / prep, which are per-snippet)."""
    import ctypes as C

    from . import _lib as L
    pb = lp["pb"]
    sub = L.CoalignProblem()
    C.memmove(C.byref(sub), C.byref(pb), C.sizeof(pb))
    sub.G = len(groups)
    row = 0
    for i, g in enumerate(groups):
        sub.d[i] = pb.d[g]
        sub.n[i] = local[g].shape[0]
        sub.row0[i] = row
        row += local[g].shape[0]
    for i in range(len(groups), L.RD_MAX_GROUPS):
        sub.n[i] = 0
        sub.row0[i] = row
    sub.n_total = row
    return sub, L.ptr_array([local[g] for g in groups])


def _sharded_merge(aligner, shard, pb, blocks, local, dev_min, s, t):
    """Per-frame partial sums over the local snippets in fp32, all-reduced; the final mean is rounded once to
    the snippet dtype exactly like the single-GPU kernel (the fp32 sum of <= ~10 fp16 values is exact)."""
    dev = aligner.device
    dt = local[0].dtype
    N, H, W, w = pb.N, pb.H, pb.W, pb.w
    acc = torch.zeros((N, H, W), dtype=torch.float32, device=dev)
    cnt = torch.zeros((N,), dtype=torch.float32, device=dev)
    mn = dev_min.to(dt)
    for g in range(pb.G):
        lo, hi = blocks[g]
        if hi == lo:
            continue
        sg = s[pb.row0[g] + lo: pb.row0[g] + hi].to(dt).view(-1, 1, 1, 1)
        tg = t[pb.row0[g] + lo: pb.row0[g] + hi].to(dt).view(-1, 1, 1, 1)
        y = ((local[g][:, :, 0] - mn) * sg + tg).float()            # same per-op rounding as depth_aligner.py:246-252
        for j in range(w):
            f0 = lo + j * pb.d[g]
            acc[f0: f0 + (hi - lo)] += y[:, j]
            cnt[f0: f0 + (hi - lo)] += 1
    shard.all_reduce_(acc, "sum")
    shard.all_reduce_(cnt, "sum")
    merged = (acc / cnt.view(-1, 1, 1)).to(dt)[:, None].contiguous()
    aligner.merged_minmax = torch.stack([merged.float().min(), merged.float().max()])
    return merged
