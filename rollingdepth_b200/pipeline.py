"""`RollingDepthPipeline` -- drop-in for the reference's pipeline on B200.

Same `__call__` / `forward` signatures, same `RollingDepthOutput` fields and the same numerics as
rollingdepth/rollingdepth_pipeline.py (cited per method); what changes is the mechanism:

  * the UNet runs with `B200XFrameAttnProcessor` (hand-written tcgen05 kernel) installed on every
    attention layer;
  * snippets are denoised `snippet_batch` at a time in one UNet call (legal: the `num_view` rearrange is
    per batch element) instead of one by one (:415-446);
  * the DDIM algebra is one fused kernel per step with host-side schedule scalars (no per-step sync);
  * co-alignment / merge / re-normalisation run in the fused kernels of `DepthAligner`;
  * with `torch.distributed` initialised (one process per GPU, `enable_sharding()`) the flattened (dilation, snippet)
    list is block-partitioned over the ranks ("snippet sharding", sharding.py); one in-place NCCL all-gather per stage
    (rgb latents, decoded snippets, refined latents, final depth) and nothing else crosses NVLink.  After the snippet
    gather every rank holds every snippet, so co-alignment / merge run the single-GPU kernels and the result is
    bit-identical to a one-GPU run; only `output_ranks` (default: rank 0) copy results to the host.

There is no CPU fallback: the pipeline must live on a CUDA device.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from os import PathLike
from typing import Dict, List, Optional, Union

import torch

from . import _lib as L
from . import ops
from .attn_processor import B200XFrameAttnProcessor
from .depth_aligner import DepthAligner
from .scheduler import DDIMSchedule
from .sharding import Shard, balanced_chunks


_HOST_POOL = None


def _host_copy_pool(world: int = 1):
    """Worker threads for the host-side result copies (torch's copy releases the GIL); `world` = number of
    processes on this box that copy results (the output ranks)."""
    global _HOST_POOL
    if _HOST_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        import os
        n = max(1, min(8, (os.cpu_count() or 1) // max(1, world)))
        _HOST_POOL = ThreadPoolExecutor(max_workers=n, thread_name_prefix="rd-hostcopy")
    return _HOST_POOL


@dataclass
class RollingDepthOutput:
    """rollingdepth_pipeline.py:44-49."""
    input_rgb: torch.Tensor
    depth_pred: torch.Tensor
    snippet_ls: Union[None, List[torch.Tensor]]
    depth_coaligned: Union[None, torch.Tensor]


class RollingDepthPipeline:
    rgb_latent_scale_factor = 0.18215
    depth_latent_scale_factor = 0.18215
    N_CHANNEL_PER_LATENT = 4

    def __init__(self, unet, vae, scheduler, text_encoder=None, tokenizer=None, snippet_batch: int = 16,
                 install_processor: bool = True, channels_last: bool = True, fused_norm: bool = True):
        self.unet = unet
        self.vae = vae
        self.scheduler = DDIMSchedule.from_scheduler(scheduler)
        self.text_encoder = text_encoder
        self.tokenizer = tokenizer
        self.empty_text_embed: torch.Tensor = None  # type: ignore
        self.snippet_batch = int(snippet_batch)
        self.shard = Shard()            # single rank unless torch.distributed is initialised (see enable_sharding)
        self.output_ranks = None        # ranks that copy results to the host (None = every rank)
        self.coalign_mode = "replicated"
        self.timings: Dict[str, float] = {}
        self.channels_last, self.fused_norm = bool(channels_last), bool(fused_norm)
        self._tuned_for = None
        self.output_mode = "fresh"      # or "pinned_views": see the output stage of forward()
        self._staging: List[torch.Tensor] = []      # pinned device -> host staging buffers, reused by every call
        self._stage_slot = 0
        self._stage_events: Dict[int, torch.cuda.Event] = {}
        self._host_copies: list = []
        self._copy_stream = None        # side stream for device -> host result copies (overlaps the remaining compute)
        self._keepalive: list = []
        if install_processor:
            self.unet.set_attn_processor(B200XFrameAttnProcessor())
        logging.debug(f"Pipeline initialized: {type(self)}")

    # ------------------------------------------------------------------ DiffusionPipeline-like surface
    def to(self, device=None, dtype=None):
        for m in (self.unet, self.vae, self.text_encoder):
            if m is not None:
                m.to(device=device, dtype=dtype)
        return self

    @property
    def device(self) -> torch.device:
        return next(self.unet.parameters()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.unet.parameters()).dtype

    def _tune_modules(self) -> None:
        """NHWC weights for cuDNN's Blackwell convolutions and the fused NHWC GroupNorm(+SiLU) kernel on every
        norm of our SD2 modules (diffusers modules are left as they are).  Idempotent; redone if the device changes."""
        key = (self.device, self.dtype)
        if self._tuned_for == key:
            return
        from .sd2.layers import BiasConv2d, GroupNormAct, LayerNormFused, ResnetBlock2D, Upsample2D
        for m in (self.unet, self.vae):
            if self.channels_last:
                m.to(memory_format=torch.channels_last)
            for sub in m.modules():
                if isinstance(sub, (GroupNormAct, BiasConv2d, ResnetBlock2D, Upsample2D)):
                    sub.fused = self.fused_norm and self.channels_last
                elif isinstance(sub, LayerNormFused):
                    sub.fused = self.fused_norm
        self._tuned_for = key

    # ------------------------------------------------------------------ optional per-stage device timing
    def _mark(self, name: str) -> None:
        """Record a CUDA event at a stage boundary when `self.stage_events` is a list (bench.py --stage-times)."""
        ev = getattr(self, "stage_events", None)
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))

    def stage_times_ms(self):
        """[(stage, ms)] between consecutive marks of the last forward (call after a synchronize)."""
        ev = getattr(self, "stage_events", None) or []
        return [(ev[i + 1][0], ev[i][1].elapsed_time(ev[i + 1][1])) for i in range(len(ev) - 1)]

    def enable_sharding(self, group=None, output_ranks=(0,), coalign_mode: str = "replicated") -> "RollingDepthPipeline":
        """Shard frames / snippets / refinement windows over the ranks of `group` (default: the world group).

        `output_ranks`: the ranks whose `forward` returns the complete `RollingDepthOutput` on the host, exactly as a
        single-GPU call would (default rank 0); the other ranks return an output whose fields are None -- copying
        ~1 GB of identical results to the host on every rank was the largest non-sharded cost at 8 GPUs.  `None` =
        every rank returns the complete result.
        `coalign_mode`: "replicated" (every rank runs the 2000-iteration optimiser on the gathered snippets; bit-identical
        everywhere, no per-iteration exchange) or "allreduce" (frames of the optimiser block-partitioned, per-iteration
        all-reduce of the per-(frame, contribution) gradient terms -- same parameters bit for bit)."""
        if coalign_mode not in ("replicated", "allreduce"):
            raise ValueError(f"unknown coalign_mode {coalign_mode!r}")
        self.shard = Shard.from_dist(group)
        self.output_ranks = None if output_ranks is None else tuple(int(r) for r in output_ranks)
        self.coalign_mode = coalign_mode
        return self

    def _is_output_rank(self) -> bool:
        return self.shard.world == 1 or self.output_ranks is None or self.shard.rank in self.output_ranks

    # ------------------------------------------------------------------ __call__ (rollingdepth_pipeline.py:78-176)
    @torch.no_grad()
    def __call__(
        self,
        input_video_path: PathLike,
        start_frame: int = 0,
        frame_count: int = 0,
        processing_res: int = 1024,
        resample_method: str = "BILINEAR",
        dilations: List[int] = [1, 25],
        cap_dilation: bool = True,
        snippet_lengths: List[int] = [3],
        init_infer_steps: List[int] = [1],
        strides: List[int] = [1],
        coalign_kwargs: Union[Dict, None] = None,
        refine_step: int = 0,
        refine_snippet_len: int = 3,
        refine_start_dilation: int = 6,
        generator: Union[torch.Generator, None] = None,
        verbose: bool = False,
        max_vae_bs: int = 4,
        unload_snippet: bool = False,
        restore_res: bool = False,
    ) -> RollingDepthOutput:
        assert processing_res >= 0
        if processing_res > 1024:
            logging.warning(f"Procssing at high-resolution ({processing_res}) may lead to suboptimal accuracy.")
        from .video_io import load_video_frames, resize_frames
        input_frames, original_res = load_video_frames(
            input_path=input_video_path, start_frame=start_frame, frame_count=frame_count,
            processing_res=processing_res, resample_method=resample_method, verbose=verbose)
        input_frames = input_frames[None]                       # "n c h w -> 1 n c h w"
        logging.info(f"{input_frames.shape[1]} frames loaded from video {input_video_path}")
        if restore_res and max(original_res) > 2048:
            logging.warning(f"Resizing back to large resolution ({list(original_res)}) may result in significant memory usage.")
        pipe_output = self.forward(
            input_frames=input_frames, dilations=dilations, cap_dilation=cap_dilation,
            snippet_lengths=snippet_lengths, init_infer_steps=init_infer_steps, strides=strides,
            coalign_kwargs=coalign_kwargs, refine_step=refine_step, refine_snippet_len=refine_snippet_len,
            refine_start_dilation=refine_start_dilation, verbose=verbose, generator=generator,
            max_vae_bs=max_vae_bs, unload_snippet=unload_snippet)
        if restore_res:
            pipe_output.input_rgb = resize_frames(pipe_output.input_rgb, list(original_res), resample_method)
            pipe_output.depth_pred = resize_frames(pipe_output.depth_pred, list(original_res), resample_method)
        return pipe_output

    def encode_empty_text(self):
        """rollingdepth_pipeline.py:178-191."""
        if self.tokenizer is None or self.text_encoder is None:
            raise RuntimeError("no text encoder/tokenizer: set `pipe.empty_text_embed` ([1, 2, cross_attention_dim])")
        text_inputs = self.tokenizer("", padding="do_not_pad", max_length=self.tokenizer.model_max_length,
                                     truncation=True, return_tensors="pt")
        ids = text_inputs.input_ids.to(self.text_encoder.device)
        self.empty_text_embed = self.text_encoder(ids)[0].to(self.dtype)

    # ------------------------------------------------------------------ forward (rollingdepth_pipeline.py:193-354)
    @torch.no_grad()
    def forward(
        self,
        input_frames: torch.Tensor,  # [1 N 3 H W]
        dilations: List[int],
        cap_dilation: bool,
        snippet_lengths: List[int],
        init_infer_steps: List[int],
        strides: List[int],
        coalign_kwargs: Union[Dict, None],
        refine_step: int,
        refine_snippet_len: int,
        refine_start_dilation: int,
        generator: Union[torch.Generator, None],
        verbose: bool,
        max_vae_bs: int,
        unload_snippet: bool,
    ) -> RollingDepthOutput:
        # ----------------- Check settings (:213-258) -----------------
        assert 1 in dilations, "dilations should include 1"
        assert len(snippet_lengths) == len(set(snippet_lengths)), f"Repeated values found in {snippet_lengths = }"
        if len(snippet_lengths) > 1:
            assert len(snippet_lengths) == len(dilations), \
                f"inconsistent lengths of window_size ({snippet_lengths}) and dilations ({dilations})"
        else:
            snippet_lengths = snippet_lengths * len(dilations)
        if len(init_infer_steps) > 1:
            assert len(init_infer_steps) == len(dilations), \
                f"inconsistent lengths of init_infer_step ({init_infer_steps}) and dilations ({dilations})"
        else:
            init_infer_steps = init_infer_steps * len(dilations)
        assert min(init_infer_steps) > 0, "Minimum inference step is 1"
        if len(strides) > 1:
            assert len(strides) == len(dilations), f"inconsistent lengths of strides ({strides}) and dilations ({dilations})"
        else:
            strides = strides * len(dilations)
        if [1] * len(dilations) != strides:
            raise NotImplementedError("Only implemented for stride 1")

        seq_len = input_frames.shape[1]
        if cap_dilation:
            for i, dilation in enumerate(dilations):        # mutates the caller's list, as the reference does (Q6)
                dilations[i] = self.cap_max_dilation(seq_len=seq_len, snippet_len=snippet_lengths[i],
                                                     dilation=dilation, verbose=verbose)
            refine_start_dilation = self.cap_max_dilation(seq_len=seq_len, snippet_len=refine_snippet_len,
                                                          dilation=refine_start_dilation, verbose=verbose)

        device = self.device
        if device.type != "cuda":
            raise L.RdError("RollingDepthPipeline (B200) must be on a CUDA device; there is no CPU fallback")
        if input_frames.shape[0] != 1:
            raise NotImplementedError("Layered inference is only implemented for B=1")
        self._tune_modules()

        # ----------------- Initial prediction -----------------
        # :263 converts on the host and then copies; copying first (async from pinned memory) and rounding to the
        # pipeline dtype on the device gives the same values (one round-to-nearest either way) without the CPU pass
        if getattr(self, "stage_events", None) is not None:
            self.stage_events = []
        self._stage_slot = 0
        self._stage_events = {}
        self._keepalive = []
        is_out = self._is_output_rank()
        self._mark("start")
        if self.shard.world > 1 and not is_out and not input_frames.is_cuda:
            # a rank that returns nothing needs only its own block of frames (for the sharded VAE encode)
            f_lo, f_hi = self.shard.block(seq_len)
            frames_dev = input_frames[:, f_lo:f_hi].to(device, non_blocking=True).to(self.dtype)
            frame_block = (f_lo, f_hi)
        else:
            frames_dev = input_frames.to(device, non_blocking=True).to(self.dtype)
            frame_block = None
        self._mark("h2d+cast")
        out_rgb = self._to_host(frames_dev.squeeze(0) / 2.0 + 0.5) if is_out else None      # :349
        rgb_latent = self.encode_rgb(frames_dev, max_batch_size=max_vae_bs, verbose=verbose, sharded=True,
                                     frame_block=frame_block, total_frames=seq_len)
        self._mark("vae_encode(+gather)")
        if self.empty_text_embed is None:
            self.encode_empty_text()
        B, seq_len, _, h, w = rgb_latent.shape

        # the same noise for every frame (:281-288)
        # (a generator living on another device, e.g. a CPU generator for reproducing a CPU run, is honoured
        #  by drawing there and copying; the reference requires generator.device == pipeline device)
        noise_dev = generator.device if generator is not None else device
        init_noise = torch.randn((B, 4, h, w), device=noise_dev, dtype=self.dtype, generator=generator).to(device)
        if self.shard.world > 1:
            self.shard.broadcast_(init_noise)
        init_noise = init_noise[:, None].expand(B, seq_len, 4, h, w)

        snippet_pred_ls = self.init_snippet_infer(
            rgb_latent=rgb_latent, init_noise=init_noise, dilations=dilations, cap_dilation=cap_dilation,
            snippet_lengths=snippet_lengths, init_infer_steps=init_infer_steps, strides=strides,
            max_vae_bs=max_vae_bs, unload_snippet=unload_snippet, verbose=verbose)
        # the decoded snippets are final: their device -> host copies overlap co-alignment (and refinement)
        out_snips = [self._to_host(snippet) for snippet in snippet_pred_ls] if is_out else None
        fresh = is_out and self.output_mode != "pinned_views"
        if fresh:       # worker threads pick these up as the transfers land, while the optimiser below runs
            out_rgb = self._detach_host(out_rgb)
            out_snips = [self._detach_host(t) for t in out_snips]
            late_shape = (seq_len, 1, *snippet_pred_ls[0].shape[-2:])
            pre_late = [self._prefault(late_shape, self.dtype) for _ in range(2)]   # depth_coaligned, depth_pred

        # ----------------- Co-alignment (:304-318) -----------------
        # sharded runs: every rank holds every snippet after the gather in init_snippet_infer, so the aligner is the
        # single-GPU one (replicated, bit-identical on every rank) unless coalign_mode == "allreduce"
        coalign_kwargs = {} if coalign_kwargs is None else coalign_kwargs
        depth_aligner = DepthAligner(verbose=verbose, device=device, **coalign_kwargs)
        depth_aligner.shard, depth_aligner.optimizer_mode = self.shard, self.coalign_mode
        depth_coaligned, scales, translations, loss_history = depth_aligner.run(
            snippet_ls=snippet_pred_ls, dilations=dilations)
        self.last_alignment = (scales, translations, loss_history)
        self.last_minmax = depth_aligner.merged_minmax          # (min, max) of the merged video before re-normalisation
        # d -= d.min(); d /= d.max(); d = d*2-1, min/max come fused out of the merge kernel
        ops.renorm_(depth_coaligned, depth_aligner.merged_minmax)
        self._mark("coalign+merge+renorm")
        out_coal = self._to_host(depth_coaligned.squeeze(0)) if is_out else None

        # ----------------- Refinement (:322-345) -----------------
        if refine_step > 0:
            depth_latent_coaligned = self.encode_rgb(
                depth_coaligned[None].expand(1, seq_len, 3, *depth_coaligned.shape[-2:]),   # "N 1 H W -> 1 N 3 H W"
                max_batch_size=max_vae_bs, verbose=verbose, sharded=True)
            depth_latent_new = self.refine(
                rgb_latent=rgb_latent, depth_latents=depth_latent_coaligned, init_noise=init_noise,
                refine_step=refine_step, snippet_len=refine_snippet_len, start_dilation=refine_start_dilation,
                verbose=verbose)
            depth_pred = self.decode_depth(depth_latent_new, max_batch_size=max_vae_bs, verbose=verbose, sharded=True)
        else:
            depth_pred = depth_coaligned
        out_pred = self._to_host(depth_pred.squeeze(0)) if is_out else None     # a separate host tensor, as :348/:352
        self._mark("refine+decode")

        # ----------------- Output (:347-354): everything to the host -----------------
        # The reference's `.cpu()` calls return fresh pageable tensors.  Page-locking ~1 GB of fresh host memory per call
        # (torch.empty(pin_memory=True)) cost 0.65 s of a 5.8 s step, so the device -> host copies go asynchronously
        # (on a side stream, as soon as each tensor is final) into pipeline-owned pinned staging buffers (allocated
        # once, reused by every call) and are then copied on the host into fresh pageable tensors (reference
        # semantics).  `output_mode = "pinned_views"` skips that host copy and returns views of the staging buffers,
        # valid until the next call of this pipeline object.
        if not is_out:
            torch.cuda.current_stream(device).synchronize()
            self._keepalive = []
            return RollingDepthOutput(input_rgb=None, depth_pred=None, snippet_ls=None, depth_coaligned=None)
        out = RollingDepthOutput(input_rgb=out_rgb, depth_pred=out_pred, snippet_ls=out_snips, depth_coaligned=out_coal)
        if fresh:
            out.depth_coaligned = self._detach_host(out.depth_coaligned, pre_late[0])
            out.depth_pred = self._detach_host(out.depth_pred, pre_late[1])
        torch.cuda.current_stream(device).synchronize()
        if self._copy_stream is not None:
            self._copy_stream.synchronize()
        for fut in self._host_copies:
            fut.result()
        self._host_copies = []
        self._keepalive = []
        self._mark("outputs_d2h")
        return out

    def _to_host(self, t: torch.Tensor) -> torch.Tensor:
        """Asynchronous device -> host copy into the next pinned staging buffer of this pipeline, issued on a side
        stream behind an event of the producing stream so that it overlaps whatever the pipeline computes next."""
        if not t.is_cuda:
            return t
        t = t.contiguous()
        self._keepalive.append(t)           # the allocator must not hand this block out again before the copy has run
        nbytes = t.numel() * t.element_size()
        slot = self._stage_slot
        self._stage_slot += 1
        if slot >= len(self._staging):
            self._staging.append(None)
        buf = self._staging[slot]
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device="cpu", pin_memory=True)
            self._staging[slot] = buf
        host = buf[:nbytes].view(t.dtype).view(t.shape)
        if self._copy_stream is None or self._copy_stream.device != t.device:
            self._copy_stream = torch.cuda.Stream(device=t.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            host.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._stage_events[host.data_ptr()] = ev
        return host

    def _host_pool(self):
        n_out = self.shard.world if self.output_ranks is None else max(1, len(self.output_ranks))
        return _host_copy_pool(n_out)

    @staticmethod
    def _row_blocks(t: torch.Tensor, n_workers: int):
        nbytes = t.numel() * t.element_size()
        rows = t.shape[0] if t.dim() > 0 else 1
        if nbytes < (16 << 20) or rows < 2:
            return [(0, rows)]
        n_blk = min(rows, 4 * n_workers)
        step = (rows + n_blk - 1) // n_blk
        return [(r0, min(r0 + step, rows)) for r0 in range(0, rows, step)]

    def _prefault(self, shape, dtype):
        """A fresh pageable host tensor whose pages are touched by the worker threads NOW (while the GPU is still busy
        with the optimiser), so that the copy of a late result into it is a plain memcpy instead of ~25 k page faults
        on the critical path at the end of the step.  Returns (tensor, {row block: future})."""
        pool = self._host_pool()
        out = torch.empty(tuple(shape), dtype=dtype, device="cpu")
        futs = {}
        for r0, r1 in self._row_blocks(out, pool._max_workers):
            futs[(r0, r1)] = pool.submit(out[r0:r1].zero_) if out.dim() > 0 else pool.submit(out.zero_)
        self._host_copies.extend(futs.values())
        return out, futs

    def _detach_host(self, t: torch.Tensor, pre=None) -> torch.Tensor:
        """Fresh pageable copy of a staging view (like the reference's .cpu()).  Nothing here blocks the calling thread:
        worker threads wait for the tensor's transfer event and then copy it in row blocks (under torchrun every rank
        runs with OMP_NUM_THREADS=1, and a single-threaded first-touch copy of ~1 GB cost 270 ms per step), so the host
        copies of the snippets run while the optimiser's launches are still being issued; forward() joins them at the end.
        `pre`: a (tensor, futures) pair from `_prefault` to copy into."""
        ev = self._stage_events.pop(t.data_ptr(), None)
        if ev is None:
            return t
        pool = self._host_pool()
        if pre is not None and tuple(pre[0].shape) == tuple(t.shape) and pre[0].dtype == t.dtype:
            out, touched = pre
        else:
            out, touched = torch.empty(t.shape, dtype=t.dtype, device="cpu"), {}

        def copy_block(dst, src, after):
            if after is not None:
                after.result()          # the block's pages have been touched (submitted earlier: cannot deadlock)
            ev.synchronize()
            dst.copy_(src)

        blocks = list(touched) if touched else self._row_blocks(out, pool._max_workers)
        for r0, r1 in blocks:
            dst, src = (out[r0:r1], t[r0:r1]) if out.dim() > 0 else (out, t)
            self._host_copies.append(pool.submit(copy_block, dst, src, touched.get((r0, r1))))
        return out

    # ------------------------------------------------------------------ initial snippet inference (:356-463)
    def init_snippet_infer(self, rgb_latent, init_noise, dilations, cap_dilation, snippet_lengths, init_infer_steps,
                           strides, max_vae_bs, unload_snippet, verbose) -> List[torch.Tensor]:
        """Returns, on EVERY rank, the decoded snippets of every dilation ([n_d, n, 1, H, W] each).

        Dilations that share (snippet length, number of denoise steps) -- all of them in every preset -- form one
        group whose snippets are flattened into a single list; that list is block-partitioned over the ranks
        (sharding.py), denoised in balanced batches, decoded in balanced chunks straight into this rank's slot of
        the gather buffer and exchanged with one in-place all-gather."""
        B, seq_len, _, h, w = rgb_latent.shape
        if 1 != B:
            raise NotImplementedError("RollingDepth is implemented for B=1")
        text = self.empty_text_embed.to(self.device)            # [1, 2, C]
        H, W = self._decoded_hw(h, w)
        per_dil: List[Optional[torch.Tensor]] = [None] * len(dilations)
        groups: Dict[tuple, List[int]] = {}
        for gi, (snippet_len, n_steps) in enumerate(zip(snippet_lengths, init_infer_steps)):
            groups.setdefault((snippet_len, n_steps), []).append(gi)
        for (snippet_len, n_steps), members in groups.items():
            timesteps = self.scheduler.set_timesteps(n_steps)
            flat: List[List[int]] = []
            spans = []
            for gi in members:
                idx_ls = self.get_snippet_indice(
                    i_step=0, timesteps=timesteps, seq_len=seq_len, snippet_len=snippet_len,
                    dilation_start=dilations[gi], dilation_end=dilations[gi], stride=strides[gi])
                spans.append((gi, len(flat), len(flat) + len(idx_ls)))
                flat.extend(idx_ls)
            n_flat = len(flat)
            lo, hi = self.shard.block(n_flat)                   # this rank's contiguous block of the flat list
            tag = ",".join(str(dilations[gi]) for gi in members)
            latents = torch.empty((hi - lo, snippet_len, 4, h, w), device=self.device, dtype=self.dtype)
            for b0, b1 in balanced_chunks(hi - lo, self.snippet_batch):
                idx = torch.tensor(flat[lo + b0: lo + b1], device=self.device)      # [b, n]
                rgb_snip = rgb_latent[0, idx]                     # [b, n, 4, h, w] (gather copy, :417)
                depth_snip = init_noise[0, idx].contiguous()
                for i_t, t in enumerate(timesteps):
                    pred = self.single_step(rgb_snip, depth_snip, t, text)
                    last = i_t == len(timesteps) - 1
                    depth_snip = self.scheduler.step(pred, t, depth_snip, out=latents[b0:b1] if last else None)
            self._mark(f"unet+ddim d={tag}")
            buf, mine = self.shard.gather_buffer(n_flat, (snippet_len, 1, H, W), self.dtype, self.device)
            if hi > lo:
                self.decode_depth(latents, max_batch_size=max_vae_bs, verbose=verbose, out=mine)
            self._mark(f"vae_decode d={tag}")
            decoded = self.shard.gather_(buf, n_flat)
            if self.shard.world > 1:
                self._mark(f"snippet_gather d={tag}")
            for gi, a, b in spans:
                per_dil[gi] = decoded[a:b].cpu() if unload_snippet else decoded[a:b]
        return per_dil      # type: ignore[return-value]

    def _decoded_hw(self, h, w):
        f = 2 ** (len(self.vae.decoder.up_blocks) - 1)
        return h * f, w * f

    @staticmethod
    def get_snippet_indice(i_step: int, timesteps, seq_len: int, snippet_len: int, dilation_start: int,
                           dilation_end: int, stride: int) -> List[List[int]]:
        """rollingdepth_pipeline.py:465-502 (the annealing expression is float-rounding sensitive, Q12)."""
        gap_start = dilation_start - 1
        gap_end = dilation_end - 1
        assert gap_start >= gap_end, f"expect gap_start > gap_end, but got {gap_start} and {gap_end}"
        assert gap_start >= 0 and gap_end >= 0
        total_step = len(timesteps)
        gap_cur = int((1 - (i_step) / total_step) * (gap_start - gap_end) + gap_end)
        total_window_size = (snippet_len - 1) * (gap_cur + 1) + 1
        i_start_ls = list(range(0, seq_len - total_window_size + 1, stride))
        if i_start_ls[-1] < seq_len - total_window_size:
            i_start_ls.append(seq_len - total_window_size)
        snippet_idx_ls = [list(range(i0, i0 + total_window_size, gap_cur + 1)) for i0 in i_start_ls]
        if not set(range(0, seq_len)) == set([x for f in snippet_idx_ls for x in f]):
            logging.warning("Not every frame is covered. Consider reducing dilation for short videos")
        return snippet_idx_ls

    @staticmethod
    def cap_max_dilation(seq_len: int, snippet_len: int, dilation: int, verbose: bool):
        """rollingdepth_pipeline.py:504-515."""
        max_allowed_gap = int(seq_len / snippet_len) - 1
        if max_allowed_gap < dilation:
            msg = f"{dilation = } is too big for {seq_len} frames. Reduced to {max_allowed_gap}"
            logging.info(msg) if verbose else logging.debug(msg)
            dilation = min(max_allowed_gap, dilation)
        return dilation

    # ------------------------------------------------------------------ refinement (:517-633)
    def refine(self, rgb_latent, depth_latents, init_noise, refine_step, snippet_len, start_dilation, verbose,
               skip_t_ratio: float = 0.5) -> torch.Tensor:
        B, seq_len, _, h, w = rgb_latent.shape
        if 1 != B:
            raise NotImplementedError("Layered inference is only implemented for B=1")
        total_scheduler_step = int(refine_step / skip_t_ratio)
        assert total_scheduler_step <= self.scheduler.num_train_timesteps, "Too many refinement steps"
        timesteps = self.scheduler.set_timesteps(total_scheduler_step)
        timesteps = timesteps[int(len(timesteps) * skip_t_ratio):]
        assert 0 < len(timesteps) < total_scheduler_step, f"invalid {skip_t_ratio = }"

        depth_latent_new = self.scheduler.add_noise(depth_latents, init_noise.contiguous(), timesteps[0])   # :548-552
        text = self.empty_text_embed.to(self.device)

        for i_step, t_current in enumerate(timesteps):
            snippet_idx_ls = self.get_snippet_indice(
                i_step=i_step, timesteps=timesteps, seq_len=seq_len, snippet_len=snippet_len,
                dilation_start=start_dilation, dilation_end=1, stride=1)
            n_win = len(snippet_idx_ls)
            g = snippet_idx_ls[0][1] - snippet_idx_ls[0][0] if snippet_len > 1 else 1
            depth_latent_old = depth_latent_new                     # every window reads the OLD latent (:605)
            lo, hi = self.shard.block(n_win)
            buf, mine = self.shard.gather_buffer(n_win, (snippet_len, 4, h, w), self.dtype, self.device)
            for b0, b1 in balanced_chunks(hi - lo, self.snippet_batch):
                idx = torch.tensor(snippet_idx_ls[lo + b0: lo + b1], device=self.device)
                rgb_in = rgb_latent[0, idx]
                depth_in = depth_latent_old[0, idx]
                pred = self.single_step(rgb_in, depth_in, t_current, text)
                self.scheduler.step(pred, t_current, depth_in, out=mine[b0:b1])
            preds = self.shard.gather_(buf, n_win)
            depth_latent_new = self._average_windows(preds, g, seq_len)[None]
        return depth_latent_new

    @staticmethod
    def _average_windows(preds: torch.Tensor, g: int, seq_len: int) -> torch.Tensor:
        """Per-frame average of the rolling windows' predictions (:585-629).  `preds` [n_win, n, C, h, w]: window k covers
        frames k + j*g (stride 1, so window starts are 0 .. n_win-1).  Accumulates per frame in the reference's order
        (windows by ascending start => j = n-1 ... 0) with the reference's in-dtype rounding after every += (:621-624),
        then divides by the integer count (:629): bit-identical to the reference's loop over windows."""
        n_win, snippet_len = preds.shape[:2]
        # `assert torch.all(count_depth_latent_new > 0)` (:628) is pure index logic: frame f is covered iff
        # f - j*g lies in [0, n_win) for some j, i.e. the w blocks [j*g, j*g + n_win) leave no gap -- checked on
        # the host instead of synchronising on `count`
        assert (snippet_len - 1) * g + n_win == seq_len, (snippet_len, g, n_win, seq_len)
        assert snippet_len == 1 and n_win == seq_len or snippet_len > 1 and n_win >= g, \
            "refinement windows do not cover every frame"
        new = torch.zeros((seq_len, *preds.shape[2:]), dtype=preds.dtype, device=preds.device)
        count = torch.zeros(seq_len, device=preds.device, dtype=torch.int32)
        for j in range(snippet_len - 1, -1, -1):
            new[j * g: j * g + n_win] += preds[:, j]
            count[j * g: j * g + n_win] += 1
        return new / count[:, None, None, None]

    # ------------------------------------------------------------------ UNet step (:635-663)
    def single_step(self, rgb_latent: torch.Tensor, depth_latent: torch.Tensor, timestep: int,
                    encoder_hidden_states: torch.Tensor) -> torch.Tensor:
        """rgb/depth latents [b, n, 4, h, w] (b snippets at once) -> model output [b, n, 4, h, w]."""
        assert depth_latent.shape == rgb_latent.shape
        b, n = rgb_latent.shape[:2]
        unet_input = torch.cat([rgb_latent.flatten(0, 1), depth_latent.flatten(0, 1)], dim=1)   # [(b n), 8, h, w]
        t_input = torch.full((b * n,), int(timestep), device=unet_input.device, dtype=torch.int64)
        ctx = encoder_hidden_states
        if ctx.shape[0] == 1 and b > 1:
            ctx = ctx.expand(b, -1, -1)          # reference code with an expanded context runs batched (Q2)
        noise_pred = self.unet(unet_input, t_input, encoder_hidden_states=ctx, num_view=n).sample
        return noise_pred.unflatten(0, (b, n))

    # ------------------------------------------------------------------ VAE wrappers (:665-740)
    def encode_rgb(self, rgb_in: torch.Tensor, max_batch_size: int, verbose: bool = False, sharded: bool = False,
                   frame_block=None, total_frames: Optional[int] = None) -> torch.Tensor:
        """:665-704.  `sharded`: this rank encodes its block of frames and the latents are all-gathered.
        `frame_block = (lo, hi)`: `rgb_in` holds only this rank's frames [lo, hi) of `total_frames`."""
        if 5 == rgb_in.dim():
            B, N, _, H, W = rgb_in.shape
        else:
            B, _, H, W = rgb_in.shape
            N = 1
        flat = rgb_in.reshape(B * N, *rgb_in.shape[-3:])
        total = B * N if frame_block is None else int(total_frames)
        lo, hi = self.shard.block(total) if sharded else (0, total)
        off = 0 if frame_block is None else frame_block[0]
        if frame_block is not None:
            assert sharded and tuple(frame_block) == (lo, hi) and B == 1, (frame_block, lo, hi)
        h, w = self._latent_hw(H, W)
        if sharded:
            buf, mine = self.shard.gather_buffer(total, (self.N_CHANNEL_PER_LATENT, h, w), self.dtype, self.device)
        else:
            buf = mine = torch.empty((total, self.N_CHANNEL_PER_LATENT, h, w), dtype=self.dtype, device=self.device)
        for i, j in balanced_chunks(hi - lo, max_batch_size):
            batch = flat[lo - off + i: lo - off + j]
            moments = self.vae.quant_conv(self.vae.encoder(batch))
            mean, _logvar = torch.chunk(moments, 2, dim=1)          # posterior mean, no sampling (Q11)
            torch.mul(mean, self.rgb_latent_scale_factor, out=mine[i:j])
        lat = self.shard.gather_(buf, total) if sharded else buf
        return lat.unflatten(0, (1, total)) if frame_block is not None else lat.unflatten(0, (B, N))

    def _latent_hw(self, H, W):
        """Spatial size of the VAE latent: every encoder level but the last pads (0,1,0,1) and applies a 3x3
        stride-2 convolution without padding (downsampling.py Downsample2D(padding=0)) = floor(size / 2)."""
        for _ in range(len(self.vae.encoder.down_blocks) - 1):
            H, W = H // 2, W // 2
        return H, W

    def decode_depth(self, depth_latent: torch.Tensor, max_batch_size: int, verbose: bool = False, sharded: bool = False,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:706-740.  `out` ([B, N, 1, H, W], contiguous): decode into it (no concatenation copy).
        `sharded`: this rank decodes its block of the B*N frames and the results are all-gathered."""
        B, N, C, H, W = depth_latent.shape
        flat = depth_latent.flatten(0, 1) / self.depth_latent_scale_factor
        Hd, Wd = self._decoded_hw(H, W)
        if sharded:
            assert out is None
            lo, hi = self.shard.block(B * N)
            buf, dst = self.shard.gather_buffer(B * N, (1, Hd, Wd), self.dtype, self.device)
        else:
            lo, hi = 0, B * N
            if out is None:
                out = torch.empty((B, N, 1, Hd, Wd), dtype=self.dtype, device=self.device)
            assert out.is_contiguous() and tuple(out.shape) == (B, N, 1, Hd, Wd), (out.shape, (B, N, 1, Hd, Wd))
            buf = dst = out.view(B * N, 1, Hd, Wd)
        for i, j in balanced_chunks(hi - lo, max_batch_size):
            stacked = self.vae.decoder(self.vae.post_quant_conv(flat[lo + i: lo + j]))
            torch.mean(stacked, dim=1, keepdim=True, out=dst[i:j])     # mean over the 3 output channels (:737)
        if sharded:
            return self.shard.gather_(buf, B * N).unflatten(0, (B, N))
        return out
