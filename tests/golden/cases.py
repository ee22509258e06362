"""Seeded inputs shared by tests/golden/make_golden.py (which runs the reference) and the parity tests."""
import zlib

import numpy as np
import torch

from osmosis_diffusion_code_b200.synthetic import synth_measurement

# create_model(**SMALL_UNET): same block types / channels-per-group as the shipped config, small spatially.
SMALL_UNET = dict(image_size=32, num_channels=256, num_res_blocks=1, channel_mult="1,2", learn_sigma=True,
                  class_cond=False, use_checkpoint=False, attention_resolutions="16", num_heads=4,
                  num_head_channels=64, num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0.0,
                  resblock_updown=True, use_fp16=False, use_new_attention_order=False, pretrain_model="osmosis")
SMALL_HW = 32

CASES = {
    "osmosis": dict(yaml="osmosis_sample_config.yaml", respacing=6, post_idx=[5, 2], step_idx=[5, 3, 0]),
    "simulation": dict(yaml="osmosis_simulation_sample_config.yaml", respacing=6, post_idx=[4], step_idx=[5, 2]),
    "haze": dict(yaml="osmosis_haze_sample_config.yaml", respacing=6, post_idx=[1], step_idx=[5, 1]),
}

_MEAS = {
    "osmosis": dict(phi_a=(1.1, 0.95, 0.95), phi_b=(0.95, 0.8, 0.8), phi_inf=(0.14, 0.29, 0.49), depth_type="gamma"),
    "simulation": dict(phi_a=(1.1, 0.95, 0.95), phi_b=(1.1, 0.95, 0.95), phi_inf=(0.2, 0.4, 0.7), depth_type="original"),
    "haze": dict(phi_a=(1.0,), phi_b=(1.0,), phi_inf=(0.14, 0.29, 0.49), depth_type="gamma"),
}


def _gen(key: str) -> torch.Generator:
    return torch.Generator().manual_seed(zlib.crc32(key.encode()))


def case_inputs(key: str):
    if key == "unet":
        g = _gen(key)
        x = torch.randn(2, 4, SMALL_HW, SMALL_HW, generator=g)
        t = torch.tensor([37, 512])
        cot = torch.randn(2, 8, SMALL_HW, SMALL_HW, generator=g)
        return x, t, cot
    kind, *rest = key.split(":")
    if kind == "meas":
        return synth_measurement(3, SMALL_HW, **_MEAS[rest[0]])
    if kind in ("x", "noise"):
        return torch.randn(1, 4, SMALL_HW, SMALL_HW, generator=_gen(key))
    raise KeyError(key)


# ---- post-processing helpers (tests/golden/make_golden_post.py) ----
POST_CASES = {"smooth64": (64, 64, 11), "noise48x80": (48, 80, 12), "ties32": (32, 32, 13), "const16": (16, 16, 14)}


def post_inputs(name: str) -> torch.Tensor:
    """Seeded depth-like [1,H,W] fp32 planes: smooth, white noise, heavily quantised (many ties), constant."""
    H, W, seed = POST_CASES[name]
    g = torch.Generator().manual_seed(seed)
    if name.startswith("smooth"):
        d = torch.nn.functional.interpolate(torch.rand(1, 1, 8, 8, generator=g), size=(H, W), mode="bilinear", align_corners=False)[0]
        return (2 * d - 1).contiguous()
    if name.startswith("noise"):
        return torch.randn(1, H, W, generator=g) * 1.7
    if name.startswith("ties"):
        return torch.round(torch.randn(1, H, W, generator=g) * 2) / 4
    return torch.full((1, H, W), 0.25)


# input-pipeline cases: (H, W, seed, kind).  Seeded uint8 images: down-scale (landscape / portrait), identity, up-scale,
# a large down-scale with long filters, and a smooth photo-like image; "grey" is a single-channel depth map.
PRE_CASES = {"land300x400": (300, 400, 21, "noise"), "port517x389": (517, 389, 22, "noise"), "same256": (256, 256, 23, "noise"),
             "up200x320": (200, 320, 24, "noise"), "big720x1280": (720, 1280, 25, "smooth"), "odd257x511": (257, 511, 26, "smooth"),
             "grey300x333": (300, 333, 27, "grey")}


def pre_inputs(name: str) -> np.ndarray:
    H, W, seed, kind = PRE_CASES[name]
    rs = np.random.RandomState(seed)
    if kind == "noise":
        return rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    ch = 1 if kind == "grey" else 3
    low = torch.from_numpy(rs.rand(1, ch, 12, 16).astype(np.float32))
    img = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)[0].clamp(0, 1)
    img = (img * 255).round().to(torch.uint8).permute(1, 2, 0).numpy()
    img = np.ascontiguousarray(img + (rs.randint(0, 3, img.shape).astype(np.uint8)))  # a little sensor noise (wraps are fine)
    return img[:, :, 0] if kind == "grey" else img


FULL_CASE = "land300x400"   # stored in full in pre_golden.npz; the other cases as every 3rd pixel + checksums


def pre_check(gold, key, got: np.ndarray, atol: float):
    """Compares `got` [3,256,256] with the golden entry `key` (full, or sub-sampled + checksum)."""
    if key in gold.files:
        return float(np.abs(got - gold[key]).max()) <= atol
    sub, sums = gold[key + "|sub3"], gold[key + "|sum"]
    ok = float(np.abs(got[:, ::3, ::3] - sub).max()) <= atol
    g64 = got.astype(np.float64)
    return ok and abs(g64.sum() - sums[0]) <= atol * got.size and abs(np.abs(g64).sum() - sums[1]) <= atol * got.size


# `ps` / rgb_guidance path (rgb_guidance_sample_config.yaml) and the `mse` loss variant of the osmosis conditioning
PS_CASE = dict(yaml="rgb_guidance_sample_config.yaml", respacing=6, step_idx=[5, 2, 0], step_seed=31)
MSE_CASE = dict(yaml="osmosis_sample_config.yaml", respacing=6, step_idx=[5, 2])


def ps_measurement():
    """A smooth RGB guide image in [-1, 1], [1,3,32,32] (the rgb_guidance demo guides the RGB channels towards an image)."""
    g = _gen("ps:meas")
    low = torch.rand(1, 3, 4, 4, generator=g)
    return (2 * torch.nn.functional.interpolate(low, size=SMALL_HW, mode="bilinear", align_corners=False) - 1).contiguous()


# BASELINE config 1 (unguided RGBD prior sampling, osmosis_utils/diffusion.py): last 6 steps of a T = 50 chain
UNCOND_CASE = dict(T=50, start_t=6, steps=6, seed=4321)


# (mean processor, variance processor, clip_denoised) combinations of the reference's registries
PROC_CASES = [("epsilon", "learned_range", True), ("start_x", "fixed_small", True), ("start_x", "learned", False),
              ("previous_x", "fixed_large", True), ("previous_x", "learned_range", False), ("epsilon", "fixed_small", False)]


def proc_inputs():
    g = _gen("proc")
    x = torch.randn(2, 4, 16, 16, generator=g)
    mo = 0.8 * x.repeat(1, 2, 1, 1) + 0.5 * torch.randn(2, 8, 16, 16, generator=g)
    return x, mo
