"""Drop-in replacement for the reference's `DepthAligner` (rollingdepth/depth_aligner.py:29-262).

Same constructor arguments and the same `run(snippet_ls, dilations)` contract; the work happens in
fused sm_100a kernels behind the C ABI (include/rd_b200.h, section 3):

  * one streaming pass for the global minimum (:78),
  * one gather for the min-shift / border crop / stride-`factor` subsample (:79-92),
  * ONE kernel per optimisation iteration (closed-form gradient of the closure at :163-214, per-frame
    targets, both normalised L1 terms, the hinge / L2 priors and the Adam update of :159, fused;
    each subsampled prediction is read once per iteration; no host synchronisation inside the loop),
  * one pass for the overlap-average merge (:231-262) that also returns min / max of the merged video
    for the pipeline's re-normalisation (rollingdepth_pipeline.py:316-318).

There is no CPU fallback: inputs are moved to `device`, which must be a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L


class DepthAligner:
    def __init__(
        self,
        device: torch.device,
        factor: int = 10,
        lmda: float = 1e-1,
        lmda2: float = 1e-1,
        lmda3: float = 1e1,
        lr: float = 1e-3,
        num_iterations: int = 2000,
        border: int = 2,
        verbose: bool = False,
        depth_loss_weight: float = 1.0,
        loss_scale=1.0,
    ):
        self.factor = factor
        self.lmda = lmda            # accepted for signature parity; unused by the reference too (:45)
        self.lmda2 = lmda2
        self.lmda3 = lmda3
        self.lr = lr
        self.num_iterations = num_iterations
        self.border = border
        self.verbose = verbose
        self.device = torch.device(device)
        self.depth_loss_weight = depth_loss_weight
        self.loss_scale = loss_scale
        # filled by run(): device tensor [2] = (min, max) of the merged video (fp32)
        self.merged_minmax: Optional[torch.Tensor] = None
        # multi-GPU (set by the pipeline, see sharding.py): with `shard.world > 1` and optimizer_mode == "allreduce"
        # the optimiser runs frame-sharded with a per-iteration all-reduce; "replicated" runs the single-GPU loop on
        # every rank (inputs are identical on every rank after the snippet all-gather, so are the results)
        self.shard = None
        self.optimizer_mode = "replicated"

    # ------------------------------------------------------------------ helpers
    def create_triplet_indices(self, sequence_length: int, gap: int, window_size: int) -> torch.Tensor:
        """depth_aligner.py:57-66."""
        d = gap + 1
        n = sequence_length - (window_size - 1) * d
        return torch.arange(max(n, 0))[:, None] + torch.arange(window_size)[None, :] * d

    def _problem(self, snippet_ls: Sequence[torch.Tensor], dilations: Sequence[int]) -> L.CoalignProblem:
        windows = [int(s.shape[1]) for s in snippet_ls]
        if len(set(windows)) != 1:
            # the reference addresses rows as i*w (:182-188), which is only valid for one window size
            raise ValueError(f"co-alignment needs one snippet length for all dilations, got {windows}")
        if len(snippet_ls) != len(dilations):
            raise ValueError("snippet_ls and dilations differ in length")
        w = windows[0]
        N = len(snippet_ls[0]) + (w - 1) * (dilations[0] - 1) + (w - 1)      # :70-76
        H, W = int(snippet_ls[0].shape[-2]), int(snippet_ls[0].shape[-1])
        pb = L.CoalignProblem()
        pb.G, pb.w, pb.N, pb.H, pb.W = len(dilations), w, N, H, W
        pb.border, pb.factor = int(self.border), int(self.factor)
        for g, d in enumerate(dilations):
            pb.d[g] = int(d)
        pb.dtype = L.dtype_code(snippet_ls[0].dtype)
        L.check(L.lib().rd_coalign_problem_init(C.byref(pb)), "rd_coalign_problem_init")
        for g, s in enumerate(snippet_ls):
            if tuple(s.shape) != (pb.n[g], w, 1, H, W):
                raise ValueError(f"snippet group {g} has shape {tuple(s.shape)}, expected {(pb.n[g], w, 1, H, W)}")
        return pb

    def _hparams(self) -> L.CoalignHParams:
        return L.CoalignHParams(self.lr, 0.5, 0.9, 1e-8, self.lmda2, self.lmda3, self.depth_loss_weight,
                                float(self.loss_scale))

    # ------------------------------------------------------------------ stages (also used by bench / tests)
    def prepare(self, snippet_ls: Sequence[torch.Tensor], dilations: Sequence[int]):
        """Global min + crop/subsample.  Returns (problem, snippets_on_device, dev_min[1], x_small)."""
        dev = self.device
        if dev.type != "cuda":
            raise L.RdError("DepthAligner runs on CUDA only (no CPU fallback)")
        dt = snippet_ls[0].dtype
        snips = [s.to(device=dev, dtype=dt).contiguous() for s in snippet_ls]
        pb = self._problem(snips, dilations)
        lib = L.lib()
        scratch = torch.empty(lib.rd_reduce_scratch_bytes(), dtype=torch.uint8, device=dev)
        dev_min = torch.empty(1, dtype=torch.float32, device=dev)
        ptrs = L.ptr_array(snips)
        st = L.stream_ptr(dev)
        L.call("rd_global_min", dev, ptrs, C.byref(pb), L.ptr(dev_min), L.ptr(scratch), st)
        x_small = torch.empty((pb.n_total, pb.w, pb.P_pad), dtype=dt, device=dev)
        L.call("rd_coalign_prep", dev, ptrs, C.byref(pb), L.ptr(dev_min), L.ptr(x_small), st)
        L.count_launches(2)
        return pb, snips, dev_min, x_small

    def new_state(self, pb: L.CoalignProblem):
        """(s, t, m_s, m_t, v_s, v_t) initialised as depth_aligner.py:150-159."""
        dev = self.device
        s = torch.ones(pb.n_total, dtype=torch.float32, device=dev)
        z = [torch.zeros(pb.n_total, dtype=torch.float32, device=dev) for _ in range(5)]
        return [s] + z

    def optimize_state(self, pb: L.CoalignProblem, x_small: torch.Tensor, state: List[torch.Tensor], step0: int,
                       n_iter: int) -> torch.Tensor:
        """Run n_iter iterations in place on `state`; returns the device history tensor [n_iter, 3]."""
        dev = self.device
        lib = L.lib()
        hp = self._hparams()
        ws_bytes = lib.rd_coalign_workspace_bytes(C.byref(pb), n_iter)
        if ws_bytes < 0:
            L.check(int(ws_bytes), "rd_coalign_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        hist = torch.empty((max(n_iter, 1), 3), dtype=torch.float32, device=dev)
        L.call("rd_coalign_optimize", dev, L.ptr(x_small), C.byref(pb), C.byref(hp), *[L.ptr(a) for a in state],
               int(step0), int(n_iter), L.ptr(hist), L.ptr(ws), ws_bytes, L.stream_ptr(dev))
        L.count_launches(n_iter + 2)      # one fused pass per iteration + final update + history reduction
        return hist[:n_iter]

    def merge(self, pb: L.CoalignProblem, snips: Sequence[torch.Tensor], dev_min: torch.Tensor, s: torch.Tensor,
              t: torch.Tensor) -> torch.Tensor:
        dev = self.device
        lib = L.lib()
        merged = torch.empty((pb.N, 1, pb.H, pb.W), dtype=snips[0].dtype, device=dev)
        self.merged_minmax = torch.empty(2, dtype=torch.float32, device=dev)
        scratch = torch.empty(lib.rd_reduce_scratch_bytes(), dtype=torch.uint8, device=dev)
        L.call("rd_merge", dev, L.ptr_array(snips), C.byref(pb), L.ptr(dev_min), L.ptr(s), L.ptr(t), L.ptr(merged),
               L.ptr(self.merged_minmax), L.ptr(scratch), L.stream_ptr(dev))
        L.count_launches(1)
        return merged

    # ------------------------------------------------------------------ reference entry point
    def run(self, snippet_ls: List[torch.Tensor], dilations: List[int]):
        """depth_aligner.py:68-120: returns (merged [N,1,H,W] in the snippet dtype, scales, translations,
        loss_history) with scales/translations = lists of fp32 [n_d,1,1] tensors per dilation and
        loss_history = list of (loss, target_min, target_max) per iteration."""
        pb, snips, dev_min, x_small = self.prepare(snippet_ls, dilations)
        covered = torch.zeros(pb.N, dtype=torch.bool)
        for g in range(pb.G):
            covered[self.create_triplet_indices(pb.N, pb.d[g] - 1, pb.w).flatten()] = True
        if not bool(covered.all()):
            logging.warning("Not every frame is covered by a snippet; uncovered frames are NaN (as in the reference)")
        state = self.new_state(pb)
        if self.shard is not None and self.shard.world > 1 and self.optimizer_mode == "allreduce":
            from .sharding import coalign_optimize_allreduce
            hist = coalign_optimize_allreduce(self, self.shard, pb, x_small, state, int(self.num_iterations))
        else:
            hist = self.optimize_state(pb, x_small, state, 0, int(self.num_iterations))
        merged = self.merge(pb, snips, dev_min, state[0], state[1])
        scales = [state[0][pb.row0[g]: pb.row0[g] + pb.n[g]].reshape(-1, 1, 1) for g in range(pb.G)]
        translations = [state[1][pb.row0[g]: pb.row0[g] + pb.n[g]].reshape(-1, 1, 1) for g in range(pb.G)]
        loss_history: List[Tuple[float, float, float]] = [tuple(r) for r in hist.cpu().tolist()]  # one sync, at the end
        return merged, scales, translations, loss_history
