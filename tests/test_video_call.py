"""`RollingDepthPipeline.__call__(input_video_path=...)` and the frame loader (rollingdepth_pipeline.py:79-176,
video_io.py:38-137): a 12-frame mp4 written with OpenCV."""
import os

import numpy as np
import pytest
import torch

import model_utils as mu


def _write_mp4(path, n=12, h=96, w=160):
    import cv2
    rng = np.random.default_rng(0)
    base = rng.integers(0, 255, size=(h // 8, w // 8, 3), dtype=np.uint8)
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (w, h))
    assert wr.isOpened()
    for i in range(n):
        frame = cv2.resize(np.roll(base, i, axis=1), (w, h), interpolation=cv2.INTER_CUBIC)
        wr.write(frame)
    wr.release()
    return n, h, w


def _decode_rgb(path):
    import cv2
    cap = cv2.VideoCapture(str(path))
    out = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        out.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
    cap.release()
    return out


def test_load_video_frames_matches_reference_semantics(tmp_path):
    """Decode -> float32 -> aspect-preserving antialiased resize to `processing_res` -> (x / 255) * 2 - 1, frame window
    [start_frame, start_frame + frame_count): checked against torchvision's resize, which the reference calls
    (video_io.py:38-68), on the same decoded frames."""
    import torchvision.transforms.functional as TF
    from torchvision.transforms import InterpolationMode
    from rollingdepth_b200.video_io import load_video_frames, resize_frames
    path = tmp_path / "clip.mp4"
    n, h, w = _write_mp4(path)
    raw = _decode_rgb(path)
    assert len(raw) == n
    frames, res = load_video_frames(path, start_frame=2, frame_count=7, processing_res=80)
    assert tuple(res) == (h, w) and tuple(frames.shape) == (7, 3, 48, 80) and frames.dtype == torch.float32
    for i in range(7):
        t = torch.from_numpy(raw[2 + i].copy()).float().permute(2, 0, 1)[None]
        f = min(80 / w, 80 / h)
        ref = TF.resize(t, [int(h * f), int(w * f)], InterpolationMode.BILINEAR, antialias=True)
        ref = (ref / 255.0) * 2.0 - 1.0
        assert torch.equal(frames[i: i + 1], ref)
    full, _ = load_video_frames(path)                 # frame_count = 0 -> everything, processing_res = 0 -> native size
    assert tuple(full.shape) == (n, 3, h, w) and float(full.min()) >= -1.0 and float(full.max()) <= 1.0
    back = resize_frames(frames, [h, w], "BILINEAR")
    assert torch.equal(back, TF.resize(frames, [h, w], InterpolationMode.BILINEAR, antialias=True))
    with pytest.raises(RuntimeError):
        load_video_frames(path, start_frame=100)


@pytest.mark.gpu
def test_call_on_a_video_file_equals_forward_on_its_frames(tmp_path):
    """`__call__` = load_video_frames + forward (+ optional restore_res): same result as `forward` on the loaded frames,
    outputs on the host, `restore_res` resizes rgb and depth back to the clip's resolution."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    from rollingdepth_b200.sd2.unet import UNet2DConditionModel
    from rollingdepth_b200.sd2.vae import AutoencoderKL
    from rollingdepth_b200.video_io import load_video_frames
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    dev = torch.device("cuda:0")
    path = tmp_path / "clip.mp4"
    n, h, w = _write_mp4(path)
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1).to(dev).half()
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(),
                                 seed=2).to(dev).half()
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=4)
    pipe.empty_text_embed = torch.randn(1, 2, 32, generator=torch.Generator().manual_seed(1)).to(dev).half()
    kw = dict(cap_dilation=True, snippet_lengths=[3], init_infer_steps=[1], strides=[1], coalign_kwargs=dict(num_iterations=30),
              refine_step=2, refine_snippet_len=3, refine_start_dilation=3, verbose=False, max_vae_bs=4, unload_snippet=False)
    dil = [1, 25]
    out = pipe(input_video_path=path, start_frame=1, frame_count=10, processing_res=80, dilations=dil,
               generator=torch.Generator(device=dev).manual_seed(0), restore_res=False, **kw)
    assert dil == [1, 2]                                    # capped in place, like the reference (Q6)
    frames, _ = load_video_frames(path, start_frame=1, frame_count=10, processing_res=80)
    ref = pipe.forward(input_frames=frames[None], dilations=[1, 25], generator=torch.Generator(device=dev).manual_seed(0), **kw)
    assert tuple(out.depth_pred.shape) == (10, 1, 48, 80) and out.depth_pred.device.type == "cpu"
    assert torch.equal(out.depth_pred, ref.depth_pred) and torch.equal(out.input_rgb, ref.input_rgb)
    assert float(out.input_rgb.min()) >= 0.0 and float(out.input_rgb.max()) <= 1.0
    assert [tuple(s.shape) for s in out.snippet_ls] == [(8, 3, 1, 48, 80), (6, 3, 1, 48, 80)]
    big = pipe(input_video_path=path, start_frame=1, frame_count=10, processing_res=80, dilations=[1, 25],
               generator=torch.Generator(device=dev).manual_seed(0), restore_res=True, **kw)
    assert tuple(big.depth_pred.shape) == (10, 1, h, w) and tuple(big.input_rgb.shape) == (10, 3, h, w)
    assert tuple(big.depth_coaligned.shape) == (10, 1, 48, 80)      # only rgb and depth_pred are restored (:168-175)
