"""Helpers shared by the fixture generator (oracle/gen_golden_pipeline.py) and the tests."""
import zlib

import torch

TINY_UNET = dict(in_channels=8, out_channels=4, down_block_types=("CrossAttnDownBlock2D", "DownBlock2D"),
                 up_block_types=("UpBlock2D", "CrossAttnUpBlock2D"), block_out_channels=(64, 128), layers_per_block=1,
                 attention_head_dim=(1, 2), cross_attention_dim=32, use_linear_projection=True, norm_num_groups=16)
TINY_VAE = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(32, 64), layers_per_block=1,
                norm_num_groups=8,
                down_block_types=("DownEncoderBlock2D", "DownEncoderBlock2D"),
                up_block_types=("UpDecoderBlock2D", "UpDecoderBlock2D"))
SD2_UNET = dict(in_channels=8, out_channels=4,
                down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
                block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, attention_head_dim=(5, 10, 20, 20),
                cross_attention_dim=1024, use_linear_projection=True, norm_num_groups=32)
SD_VAE = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
              layers_per_block=2, norm_num_groups=32,
              down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4)
DDIM_CFG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="v_prediction",
                timestep_spacing="trailing")


@torch.no_grad()
def deterministic_init_(module: torch.nn.Module, seed: int = 0, gain: float = 1.0) -> torch.nn.Module:
    """Fill every parameter from a generator keyed by the parameter NAME, so two implementations with
    identical state-dict keys get identical weights without storing them."""
    for name, p in sorted(module.state_dict().items()):
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) % (2 ** 31))
        if p.dim() <= 1:
            if name.endswith("weight"):        # norm scales
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = 0.02 * torch.randn(p.shape, generator=g)
        else:
            fan_in = p[0].numel()
            v = torch.randn(p.shape, generator=g) * (gain / fan_in ** 0.5)
        p.copy_(v.to(p.dtype))
    return module


def synth_video(n, h, w, seed=0):
    """Smooth synthetic clip in [-1, 1], float32 [1, n, 3, h, w] (same generator as bench.py's `synth_video`)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(n, 3, h // 16 + 2, w // 16 + 2, generator=g)
    vid = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    drift = torch.linspace(0, 1, n)[:, None, None, None] * 0.2
    return ((vid + drift).clamp(0, 1) * 2 - 1)[None].contiguous()
