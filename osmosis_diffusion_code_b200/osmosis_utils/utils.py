"""Host-side helpers of the sampling path: config loading and the sampling-pattern phase logic.

Mirrors, from the reference `osmosis_utils/utils.py`: `load_yaml` / `arguments_from_file` (:357-360, :466-476),
`str2bool` (:384-395), `get_depth_value` (:529-541), `is_freeze_phi` (:571-590), `set_alternate_length`
(:595-630).  Everything image / visualisation related in that file is out of scope.
"""
from __future__ import annotations

import argparse

import numpy as np
import yaml


def load_yaml(file_path: str) -> dict:
    with open(file_path) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def arguments_from_file(config_file_path: str) -> argparse.Namespace:
    """Top-level YAML keys become attributes; nested mappings stay dicts (they are splatted into factories)."""
    ns = argparse.Namespace()
    for key, value in load_yaml(config_file_path).items():
        setattr(ns, key, value)
    return ns


def str2bool(v):
    if isinstance(v, bool):
        return v
    s = v.lower()
    if s in ("yes", "true", "t", "y", "1"):
        return True
    if s in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected")


def get_depth_value(value_raw, **kwargs):
    if isinstance(value_raw, float):
        return value_raw
    if isinstance(value_raw, int):
        return float(value_raw)
    if isinstance(value_raw, str):
        return np.fromstring(value_raw, dtype=float, sep=",")
    if isinstance(value_raw, (np.ndarray, np.generic)):
        return value_raw
    raise NotImplementedError


def _outside(sample_pattern, lo_key, hi_key, time_index, num_timesteps):
    return time_index > sample_pattern[hi_key] * num_timesteps or time_index < sample_pattern[lo_key] * num_timesteps


def is_freeze_phi(sample_pattern, time_index, num_timesteps):
    """phi is frozen outside the guidance window and outside [update_end, update_start] * T."""
    if sample_pattern is None or sample_pattern["pattern"] == "original":
        return False
    if _outside(sample_pattern, "stop_guidance", "start_guidance", time_index, num_timesteps):
        return True
    return _outside(sample_pattern, "update_end", "update_start", time_index, num_timesteps)


def set_alternate_length(sample_pattern, time_index, num_timesteps):
    """Number of x / phi alternations at this step (gibbsDDRM's M); 1 outside [s_end, s_start] * T."""
    if sample_pattern is None or sample_pattern["pattern"] == "original":
        return 1
    assert sample_pattern["update_start"] > sample_pattern["update_end"]
    assert sample_pattern["s_start"] > sample_pattern["s_end"]
    if sample_pattern["local_M"] > 1:
        assert sample_pattern["update_start"] >= sample_pattern["s_start"]
        assert sample_pattern["s_end"] >= sample_pattern["update_end"]
    for lo, hi in (("stop_guidance", "start_guidance"), ("update_end", "update_start"), ("s_end", "s_start")):
        if _outside(sample_pattern, lo, hi, time_index, num_timesteps):
            return 1
    return sample_pattern["local_M"]


# ---------------------------------------------------------------------------------------------------------------------
# Post-processing of finished samples on the device (reference: osmosis_utils/utils.py:46-114, 748-763 and the per-image
# CPU block of osmosis_sampling.py:207-292).  Same names / argument meaning as the reference; tensors stay on the GPU and a
# leading batch dimension is processed in one launch (a batch of B == B separate reference calls).
# ---------------------------------------------------------------------------------------------------------------------
def _planes(img):
    """[C,H,W] -> one plane of C*H*W elements (the reference reduces over the whole tensor); [B,...] -> B planes."""
    import torch
    from .. import lib as _lib
    if not img.is_cuda:
        raise _lib.OsmError("post-processing runs on CUDA tensors only (no CPU fallback)")
    if img.dim() == 3:
        return img.contiguous().float().view(1, -1)
    if img.dim() == 4:
        return img.contiguous().float().view(img.shape[0], -1)
    raise NotImplementedError


def min_max_norm_range(img, vmin=0, vmax=1, is_uint8=False):
    """utils.py:46-76: affine map of [min, max] (per image) to [vmin, vmax]; zeros if the image is constant."""
    return min_max_norm_range_percentile(img, vmin=vmin, vmax=vmax, percent_low=0.0, percent_high=1.0, is_uint8=is_uint8)


def min_max_norm_range_percentile(img, vmin=0, vmax=1, percent_low=0., percent_high=1., is_uint8=False):
    """utils.py:79-114: clip to the (percent_low, percent_high) quantiles (torch.quantile, linear interpolation), then
    min-max normalise.  The quantiles are exact order statistics (radix select on the device), bit-identical to
    torch.quantile on the same values."""
    import torch
    from .. import lib as _lib
    p = _planes(img)
    out = torch.empty_like(p)
    L = _lib.load()
    _lib.check(L.osm_minmax_percentile(_lib.ptr(p), _lib.ptr(out), p.shape[0], p.shape[1], float(percent_low), float(percent_high),
                                       float(vmin), float(vmax), _lib.stream()))
    out = out.view(img.shape)
    if is_uint8:
        out = (255 * out).to(torch.uint8)
    return out


_VIRIDIS_POLY = ((0.2777273272234177, 0.005407344544966578, 0.3340998053353061),
                 (0.1050930431085774, 1.404613529898575, 1.384590162594685),
                 (-0.3308618287255563, 0.214847559468213, 0.09509516302823659),
                 (-4.634230498983486, -5.799100973351585, -19.33244095627987),
                 (6.228269936347081, 14.17993336680509, 56.69055260068105),
                 (4.776384997670288, -13.74514537774601, -65.35303263337234),
                 (-5.435455855934631, 4.645852612178535, 26.3124352495832))


def colormap_table(colormap="viridis"):
    """[256,3] float32 RGB table of a matplotlib colormap.  Taken from matplotlib when it is importable (what the reference
    uses, utils.py:749); otherwise - matplotlib is not a dependency of this package - viridis falls back to a degree-6
    polynomial fit of the table (max deviation ~0.01), any other name raises."""
    import numpy as np
    try:
        import matplotlib.pyplot as plt
        cm = plt.get_cmap(colormap)
        return np.asarray(cm(np.arange(256)))[:, :3].astype(np.float32)
    except Exception:
        if colormap != "viridis":
            raise NotImplementedError(f"colormap '{colormap}' needs matplotlib")
        t = np.arange(256, dtype=np.float64)[:, None] / 255.0
        c = np.zeros((256, 3))
        for coef in reversed(_VIRIDIS_POLY):
            c = c * t + np.asarray(coef)[None, :]
        return np.clip(c, 0.0, 1.0).astype(np.float32)


def depth_tensor_to_color_image(tensor_image, colormap="viridis", lut=None):
    """utils.py:748-763: colour a [0,1] depth map.  Accepts [H,W], [1,H,W], [1,1,H,W] like the reference (-> [3,H,W]) and
    also a batch [B,1,H,W] (-> [B,3,H,W]).  `lut`: optional [256,3] table (numpy / tensor) overriding `colormap`."""
    import torch
    from .. import lib as _lib
    if not tensor_image.is_cuda:
        raise _lib.OsmError("post-processing runs on CUDA tensors only (no CPU fallback)")
    t = tensor_image
    batched = t.dim() == 4 and t.shape[0] > 1
    if t.dim() == 4 and not batched:
        t = t.squeeze()
    if t.dim() == 3 and not batched:
        t = t[0]
    if batched:
        t = t[:, 0]
    else:
        assert t.dim() == 2
        t = t[None]
    t = t.contiguous().float()
    table = torch.as_tensor(colormap_table(colormap) if lut is None else lut, dtype=torch.float32).to(t.device).contiguous()
    B, H, W = t.shape
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=t.device)
    L = _lib.load()
    _lib.check(L.osm_colormap(_lib.ptr(t), _lib.ptr(table), _lib.ptr(out), B, H * W, _lib.stream()))
    return out if batched else out[0]


def postprocess_samples(operator, pred_xstart, measurement):
    """The per-image block after p_sample_loop in osmosis_sampling.py:207-292, batched on the device.

    operator: the measurement operator used for sampling (holds phi), pred_xstart [B,4,H,W], measurement [B,3,H,W] in
    [-1,1] (ref_img).  Returns a dict of device tensors: sample_rgb_01_clip, degraded_image, sample_rgb_recon [B,3,H,W],
    norm_loss [B], sample_depth_mm / sample_depth_vis_pmm [B,1,H,W] and sample_depth_vis_pmm_color [B,3,H,W]."""
    import torch
    from .. import lib as _lib
    x0 = pred_xstart.to(measurement.device).contiguous().float()
    y = measurement.contiguous().float()
    if not y.is_cuda:
        raise _lib.OsmError("post-processing runs on CUDA tensors only (no CPU fallback)")
    B, _, H, W = x0.shape
    if int(operator.phi.shape[0]) != B or int(y.shape[0]) != B:
        raise ValueError(f"postprocess_samples: {B} samples, {int(y.shape[0])} measurements, phi for {int(operator.phi.shape[0])} images")
    rgb = torch.empty(B, 3, H, W, dtype=torch.float32, device=y.device)
    deg, rec = torch.empty_like(rgb), torch.empty_like(rgb)
    norm = torch.empty(B, dtype=torch.float32, device=y.device)
    L = _lib.load()
    dv = (_lib.C.c_float * 3)(*operator.depth_val)
    from ..guided_diffusion.condition_methods import OP_KIND
    _lib.check(L.osm_postprocess(OP_KIND[operator.kind_name], operator.depth_kind, dv, _lib.ptr(x0), _lib.ptr(y), _lib.ptr(operator.phi),
                                 _lib.ptr(rgb), _lib.ptr(deg), _lib.ptr(rec), _lib.ptr(norm), B, H * W, _lib.stream()))
    depth = x0[:, 3:4].contiguous()
    depth_mm = min_max_norm_range(depth)
    depth_pmm = min_max_norm_range_percentile(depth, vmin=0, vmax=1, percent_low=0.03, percent_high=0.99)
    return dict(sample_rgb_01_clip=rgb, degraded_image=deg, sample_rgb_recon=rec, norm_loss=norm, sample_depth_mm=depth_mm,
                sample_depth_vis_pmm=depth_pmm,
                sample_depth_vis_pmm_color=depth_tensor_to_color_image(depth_pmm) if B > 1 else depth_tensor_to_color_image(depth_pmm)[None])
