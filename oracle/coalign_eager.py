"""Oracle #2 for depth co-alignment: the reference's OWN formulation, device-agnostic (TEST INFRASTRUCTURE).

oracle/coalign.py restates the optimisation with the closed-form gradient the CUDA kernel implements.  This file
restates it the way /root/reference/rollingdepth/depth_aligner.py:123-229 actually computes it -- dense
[G*w, N, P] scatter buffers, autograd through the closure, torch.optim.Adam -- so that

  * tests can check closed form == autograd formulation on the same inputs (tests/test_oracle_vs_golden.py), and
  * bench.py can time "what a user of the reference gets on this GPU" (`gpu_reference`): ~40 small kernels and three
    host synchronisations per iteration (`loss.item()`, `summ.min().item()`, `summ.max().item()`, :213), three dense
    temporaries re-allocated per iteration.

Runs on whatever device the snippets live on.  Never imported by the product.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .coalign import create_triplet_indices, sequence_length_of


def optimize_autograd(x_ls: Sequence[torch.Tensor], dilations: Sequence[int], N: int, num_iterations: int = 2000,
                      lr: float = 1e-3, lmda2: float = 1e-1, lmda3: float = 1e1, depth_loss_weight: float = 1.0,
                      loss_scale: float = 1.0):
    """x_ls[i]: [n_i, w, P] subsampled, min-shifted predictions (any float dtype; promoted by the fp32 parameters).
    Returns (scales [n_i,1,1] list, translations list, history [(loss, target_min, target_max)])."""
    dev = x_ls[0].device
    P = x_ls[0].shape[-1]
    windows = [x.shape[1] for x in x_ls]
    rows_total = sum(windows)
    idx_ls = [create_triplet_indices(N, d - 1, w).to(dev) for d, w in zip(dilations, windows)]      # [n_i, w]
    scales = [torch.ones(x.shape[0], 1, 1, device=dev, requires_grad=True) for x in x_ls]            # :150-153
    shifts = [torch.zeros(x.shape[0], 1, 1, device=dev, requires_grad=True) for x in x_ls]           # :154-157
    opt = torch.optim.Adam(scales + shifts, lr=lr, betas=(0.5, 0.9))                                 # :159
    history = []

    def closure():
        dense = torch.zeros(rows_total, N, P, device=dev)          # "M"       (:168-170)
        dense_inv = torch.zeros(rows_total, N, P, device=dev)      # "M_depth" (:171-173)
        present = torch.zeros(rows_total, N, P, device=dev)        # "B"       (:174-176)
        for i, (x, s, t, idx, w) in enumerate(zip(x_ls, scales, shifts, idx_ls, windows)):
            y = x * s + t                                           # :163-166, fp32 by promotion
            rows = torch.arange(i * w, (i + 1) * w, device=dev)[:, None]        # reference row rule i*w (:182)
            cols = idx.T                                            # [w, n_i]
            dense_inv[rows, cols] = y.clip(1e-3).permute(1, 0, 2) ** -1
            dense[rows, cols] = y.permute(1, 0, 2)
            present[rows, cols] = 1
        count = present.sum(0)
        mean_y, mean_inv = dense.sum(0) / count, dense_inv.sum(0) / count      # :190-191
        with torch.no_grad():
            target, target_inv = mean_y.clone().detach(), mean_inv.clone().detach()
            a = target.abs().mean(-1, keepdim=True)
            b = target_inv.abs().mean(-1, keepdim=True)
        l_y = torch.abs((dense - target) * present / a).mean()      # over ALL entries, empty ones included (:200)
        l_inv = torch.abs((dense_inv - target_inv) * present / b).mean()
        data = (l_y + depth_loss_weight * l_inv).mean()
        prior = sum(lmda2 * (torch.max(torch.tensor(0.0), 1 - s) ** 2).mean() + lmda3 * (t ** 2).mean()
                    for s, t in zip(scales, shifts))
        loss = loss_scale * data + prior
        loss.backward()
        history.append((loss.item(), mean_y.min().item(), mean_y.max().item()))     # three host syncs (:213)
        return loss

    for _ in range(num_iterations):                                # :217-222
        opt.zero_grad()
        opt.step(closure)
    return scales, shifts, history


class EagerDepthAligner:
    """`DepthAligner.run` (depth_aligner.py:68-120) with eager torch ops on the snippets' device: global min, crop,
    subsample, the autograd optimiser above and the per-frame masked-gather merge (:231-262)."""

    def __init__(self, device=None, factor=10, lmda=1e-1, lmda2=1e-1, lmda3=1e1, lr=1e-3, num_iterations=2000, border=2,
                 verbose=False, depth_loss_weight=1.0, loss_scale=1.0):
        self.device, self.factor, self.lmda2, self.lmda3, self.lr = device, factor, lmda2, lmda3, lr
        self.num_iterations, self.border = num_iterations, border
        self.depth_loss_weight, self.loss_scale = depth_loss_weight, loss_scale

    def run(self, snippet_ls: List[torch.Tensor], dilations: List[int]):
        dev = self.device if self.device is not None else snippet_ls[0].device
        N = sequence_length_of(snippet_ls, dilations)
        mn = min([s.min() for s in snippet_ls])
        shifted = [s - mn for s in snippet_ls]
        b, f = self.border, self.factor
        small = [s[:, :, :, b:-b, b:-b].to(dev)[:, :, :, ::f, ::f] for s in shifted]
        small = [s.reshape(s.shape[0], s.shape[1], -1) for s in small]
        scales, shifts, hist = optimize_autograd(small, dilations, N, self.num_iterations, self.lr, self.lmda2, self.lmda3,
                                                 self.depth_loss_weight, self.loss_scale)
        merged = self.merge(shifted, scales, shifts, dilations, N, dev)
        return merged, scales, shifts, hist

    @staticmethod
    def merge(shifted, scales, shifts, dilations, N, dev):
        shifted = [a.to(dev) for a in shifted]
        dt = shifted[0].dtype
        scaled = [x * s[:, None, None].to(dt).to(dev) + t[:, None, None].to(dt).to(dev)
                  for x, s, t in zip(shifted, scales, shifts)]
        idx_ls = [create_triplet_indices(N, d - 1, x.shape[1]).to(dev) for d, x in zip(dilations, shifted)]
        frames = []
        for f in range(N):                                         # N * G masked gathers, each with a nonzero sync (:256-260)
            parts = [a[idx == f] for a, idx in zip(scaled, idx_ls)]
            frames.append(torch.cat(parts).mean(0))
        return torch.cat(frames)[:, None]
