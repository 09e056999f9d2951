"""Oracle for depth co-alignment (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates /root/reference/rollingdepth/depth_aligner.py with torch CPU tensors.
The 2000-step optimisation is restated with the CLOSED-FORM gradient of the
reference's closure (no autograd), which is what the CUDA kernels implement;
``tests/test_oracle_vs_golden.py`` pins it against fixtures produced by the
reference's own autograd closure and Adam (oracle/gen_golden.py).

Notation: groups i (one per dilation d_i), uniform window w,
n_i = N - (w-1) d_i snippets per group, snippet k of group i covers frames
f = k + j d_i, j = 0..w-1.  x[i] has shape [n_i, w, P].
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch


# --------------------------------------------------------------------------- indices
def create_triplet_indices(sequence_length: int, gap: int, window_size: int) -> torch.Tensor:
    """depth_aligner.py:57-66 -- indices[k, j] = k + j*(gap+1), int64 [n, w]."""
    d = gap + 1
    n = sequence_length - (window_size - 1) * d
    k = torch.arange(max(n, 0), dtype=torch.int64)[:, None]
    j = torch.arange(window_size, dtype=torch.int64)[None, :]
    return k + j * d


def sequence_length_of(snippet_ls: Sequence[torch.Tensor], dilations: Sequence[int]) -> int:
    """depth_aligner.py:70-76 -- N = n_0 + (w_0-1)*gap_0 + (w_0-1)."""
    w0 = snippet_ls[0].shape[1]
    return len(snippet_ls[0]) + (w0 - 1) * (dilations[0] - 1) + (w0 - 1)


# --------------------------------------------------------------------------- prep
def prep(snippet_ls: Sequence[torch.Tensor], border: int = 2, factor: int = 10):
    """depth_aligner.py:78-92.

    Returns (mn, shifted_ls, small_ls): the global minimum (0-dim, snippet dtype),
    the min-shifted full-resolution snippets (arithmetic in the snippet dtype, as
    the reference does: fp16 on the `fast` preset) and the border-cropped,
    stride-`factor` subsampled copies reshaped to [n, w, H'*W'].
    """
    mn = min([s.min() for s in snippet_ls])
    shifted = [s - mn for s in snippet_ls]
    small = []
    for s in shifted:
        c = s[:, :, :, border:-border, border:-border]
        c = c[:, :, :, ::factor, ::factor]
        small.append(c.reshape(c.shape[0], c.shape[1], -1).contiguous())
    return mn, shifted, small


# --------------------------------------------------------------------------- loss + gradient
def loss_and_grad(
    x_ls: Sequence[torch.Tensor],
    s_ls: Sequence[torch.Tensor],
    t_ls: Sequence[torch.Tensor],
    dilations: Sequence[int],
    N: int,
    lmda2: float = 1e-1,
    lmda3: float = 1e1,
    depth_loss_weight: float = 1.0,
    loss_scale: float = 1.0,
    dtype: torch.dtype = torch.float32,
):
    """Closed form of the closure at depth_aligner.py:163-214.

    x_ls[i]: [n_i, w, P] (any float dtype; promoted like `x * s + t` with fp32
    params, :165).  s_ls[i], t_ls[i]: [n_i] (or [n_i,1,1]).  Returns
    (loss, tmin, tmax, gs_ls, gt_ls) with python floats / [n_i] tensors.
    """
    G = len(x_ls)
    w = x_ls[0].shape[1]
    P = x_ls[0].shape[2]
    R = G * w
    Z = float(R) * float(N) * float(P)  # `.mean()` divides by ALL entries (:200-201)

    sum_y = torch.zeros(N, P, dtype=dtype)
    sum_inv = torch.zeros(N, P, dtype=dtype)
    cnt = torch.zeros(N, dtype=dtype)
    ys, invs, idxs, xs = [], [], [], []
    for i in range(G):
        x = x_ls[i].to(dtype)
        s = s_ls[i].reshape(-1, 1, 1).to(dtype)
        t = t_ls[i].reshape(-1, 1, 1).to(dtype)
        y = x * s + t                                        # :165
        inv = 1.0 / y.clamp(min=1e-3)                        # :183-184
        idx = create_triplet_indices(N, dilations[i] - 1, w)  # [n_i, w]
        assert idx.shape[0] == x.shape[0], (idx.shape, x.shape)
        for j in range(w):                                   # row order i*w + j (:182-188)
            sum_y[idx[:, j]] += y[:, j]
            sum_inv[idx[:, j]] += inv[:, j]
            cnt[idx[:, j]] += 1
        ys.append(y); invs.append(inv); idxs.append(idx); xs.append(x)

    target = sum_y / cnt[:, None]                            # :190
    tdep = sum_inv / cnt[:, None]                            # :191
    a = target.abs().mean(-1)                                # :197  [N]
    b = tdep.abs().mean(-1)                                  # :198

    l1 = torch.zeros((), dtype=dtype)
    l2 = torch.zeros((), dtype=dtype)
    gs_ls, gt_ls = [], []
    soft = torch.zeros((), dtype=dtype)
    for i in range(G):
        y, inv, idx, x = ys[i], invs[i], idxs[i], xs[i]
        tg = target[idx]                                     # [n_i, w, P]
        td = tdep[idx]
        ia = (1.0 / a)[idx][..., None]                       # [n_i, w, 1]
        ib = (1.0 / b)[idx][..., None]
        r1 = y - tg
        r2 = inv - td
        l1 = l1 + (r1.abs() * ia).sum()
        l2 = l2 + (r2.abs() * ib).sum()
        live = (y >= 1e-3).to(dtype)                         # clip passes gradient only when not clipped
        gy = (loss_scale / Z) * (
            torch.sign(r1) * ia - depth_loss_weight * torch.sign(r2) * inv * inv * live * ib
        )
        s = s_ls[i].reshape(-1).to(dtype)
        t = t_ls[i].reshape(-1).to(dtype)
        n_i = float(s.numel())
        gs = (gy * x).sum((1, 2)) - 2.0 * lmda2 * torch.relu(1.0 - s) / n_i   # :205-209
        gt = gy.sum((1, 2)) + 2.0 * lmda3 * t / n_i
        gs_ls.append(gs); gt_ls.append(gt)
        soft = soft + lmda2 * (torch.relu(1.0 - s) ** 2).mean() + lmda3 * (t ** 2).mean()
    loss = loss_scale * (l1 / Z + depth_loss_weight * l2 / Z) + soft          # :203, :211
    return float(loss), float(target.min()), float(target.max()), gs_ls, gt_ls


# --------------------------------------------------------------------------- Adam
class AdamState:
    """torch.optim.Adam(lr, betas=(0.5, 0.9), eps=1e-8) as used at depth_aligner.py:159."""

    def __init__(self, shapes: Sequence[int], dtype=torch.float32):
        self.m = [torch.zeros(n, dtype=dtype) for n in shapes]
        self.v = [torch.zeros(n, dtype=dtype) for n in shapes]
        self.step = 0


def adam_update(params, grads, state: AdamState, lr=1e-3, beta1=0.5, beta2=0.9, eps=1e-8):
    """One torch-semantics Adam step; params updated in place.

    m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ;
    theta <- theta - lr/(1-b1^k) * m / (sqrt(v)/sqrt(1-b2^k) + eps)
    """
    state.step += 1
    k = state.step
    bc1 = 1.0 - beta1 ** k
    bc2_sqrt = math.sqrt(1.0 - beta2 ** k)
    step_size = lr / bc1
    for p, g, m, v in zip(params, grads, state.m, state.v):
        m.mul_(beta1).add_(g, alpha=1.0 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def optimize(
    x_ls, dilations, N, num_iterations=2000, lr=1e-3, lmda2=1e-1, lmda3=1e1,
    depth_loss_weight=1.0, loss_scale=1.0, dtype=torch.float32,
    init=None,
):
    """depth_aligner.py:123-229.  Returns (scales, translations, history).

    `init` = optional (s_ls, t_ls, m_s, m_t, v_s, v_t, step) for teacher-forced runs.
    """
    ns = [x.shape[0] for x in x_ls]
    if init is None:
        s_ls = [torch.ones(n, dtype=dtype) for n in ns]
        t_ls = [torch.zeros(n, dtype=dtype) for n in ns]
        st = AdamState(ns + ns, dtype)
    else:
        s_ls = [a.clone().to(dtype).reshape(-1) for a in init[0]]
        t_ls = [a.clone().to(dtype).reshape(-1) for a in init[1]]
        st = AdamState(ns + ns, dtype)
        st.m = [a.clone().to(dtype).reshape(-1) for a in (list(init[2]) + list(init[3]))]
        st.v = [a.clone().to(dtype).reshape(-1) for a in (list(init[4]) + list(init[5]))]
        st.step = int(init[6])
    hist = []
    for _ in range(num_iterations):
        loss, tmin, tmax, gs, gt = loss_and_grad(
            x_ls, s_ls, t_ls, dilations, N, lmda2, lmda3, depth_loss_weight, loss_scale, dtype
        )
        hist.append((loss, tmin, tmax))                      # before the update (:213)
        adam_update(s_ls + t_ls, gs + gt, st, lr=lr)
    return s_ls, t_ls, hist, st


# --------------------------------------------------------------------------- merge
def merge(shifted_ls, s_ls, t_ls, dilations, N):
    """depth_aligner.py:231-262: per-frame mean of x*s+t in the SNIPPET dtype.

    shifted_ls[i]: [n_i, w, 1, H, W] already min-shifted.  Returns [N, 1, H, W].
    """
    dtype = shifted_ls[0].dtype
    w = shifted_ls[0].shape[1]
    scaled = [
        x * s.reshape(-1, 1, 1, 1, 1).to(dtype) + t.reshape(-1, 1, 1, 1, 1).to(dtype)   # :246-252
        for x, s, t in zip(shifted_ls, s_ls, t_ls)
    ]
    idxs = [create_triplet_indices(N, d - 1, w) for d in dilations]
    out = []
    for f in range(N):
        parts = [a[idx == f] for a, idx in zip(scaled, idxs)]                            # :258
        out.append(torch.cat(parts).mean(0))                                             # :260
    return torch.cat(out)[:, None]


def renorm(depth: torch.Tensor) -> torch.Tensor:
    """rollingdepth_pipeline.py:316-318 (sequential in-dtype arithmetic)."""
    d = depth - depth.min()
    d = d / d.max()
    return d * 2.0 - 1.0


class DepthAlignerRef:
    """Same ctor/run contract as depth_aligner.py:29-120, closed-form inside."""

    def __init__(self, device=None, factor=10, lmda=1e-1, lmda2=1e-1, lmda3=1e1, lr=1e-3,
                 num_iterations=2000, border=2, verbose=False, depth_loss_weight=1.0, loss_scale=1.0):
        self.factor, self.lmda2, self.lmda3, self.lr = factor, lmda2, lmda3, lr
        self.num_iterations, self.border = num_iterations, border
        self.depth_loss_weight, self.loss_scale = depth_loss_weight, loss_scale

    def run(self, snippet_ls: List[torch.Tensor], dilations: List[int]):
        snippet_ls = [s.detach().cpu() for s in snippet_ls]
        N = sequence_length_of(snippet_ls, dilations)
        _, shifted, small = prep(snippet_ls, self.border, self.factor)
        s_ls, t_ls, hist, _ = optimize(
            small, dilations, N, self.num_iterations, self.lr, self.lmda2, self.lmda3,
            self.depth_loss_weight, self.loss_scale,
        )
        merged = merge(shifted, s_ls, t_ls, dilations, N)
        return merged, [s.reshape(-1, 1, 1) for s in s_ls], [t.reshape(-1, 1, 1) for t in t_ls], hist
