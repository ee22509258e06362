"""Unconditional RGBD ancestral sampling (BASELINE config 1, `RGBD_prior_sampling.py`).

Mirrors the reference `osmosis_utils/diffusion.py`: `GaussianDiffusion(T, schedule)` (:19-46) and `.inverse`
(:59-130).  Quirks kept on purpose: timesteps are 1-indexed floats, the linear
schedule is NOT rescaled by 1000/T, `steps < T` gives a truncated (not respaced) chain, x_T is drawn on the
CPU and moved, z is drawn for every t > 1.  The process record (:98-127) is reproduced from device snapshots; the update x <- (x - (1-a)/sqrt(1-abar) eps)/sqrt(a) + sqrt(beta~) z is one kernel
(osm_ddpm_uncond_update).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import lib as _lib


class GaussianDiffusion:
    def __init__(self, T, schedule):
        self.T = T
        if schedule == "linear":
            self.beta = np.linspace(1e-4, 2e-2, T)
        elif schedule == "cosine":
            f = lambda t: np.cos(math.pi * 0.5 * (t / T + 0.008) / 1.008) ** 2
            abar = f(np.arange(0, T + 1, 1)) / f(0)
            self.beta = np.clip(1 - (abar[1:] / abar[:-1]), None, 0.999)
        else:
            raise NotImplementedError(f"unknown schedule: {schedule}")
        self.betabar = np.cumprod(self.beta)
        self.alpha = 1 - self.beta
        self.alphabar = np.cumprod(self.alpha)

    def step_coefficients(self, t):
        """(c_x, c_eps, c_z) of x <- c_x (x - c_eps eps) + c_z z at 1-indexed step t, rounded like the reference
        (numpy float64 scalars multiplying fp32 tensors act as fp32 scalars)."""
        at, atbar = self.alpha[t - 1], self.alphabar[t - 1]
        beta_tilde = self.beta[t - 1] * (1 - self.alphabar[t - 2]) / (1 - atbar) if t > 1 else 0.0
        return float(1 / np.sqrt(at)), float((1 - at) / np.sqrt(1 - atbar)), float(np.sqrt(beta_tilde))

    def inverse(self, net, shape=(1, 64, 64), image_channels=3, steps=None, x=None, start_t=None, device="cpu", **kwargs):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.OsmError("inverse() runs on CUDA devices only (no CPU fallback)")
        if x is None:
            x = torch.randn((1,) + tuple(shape)).to(device)
        start_t = self.T if start_t is None else start_t
        steps = self.T if steps is None else steps
        x = x.contiguous().float()
        B, Cc, H, W = x.shape
        L = _lib.load()
        record_process = kwargs.get("record_process", False)
        record_every = kwargs.get("record_every", 200)
        save_path = kwargs.get("save_path", None)
        image_idx = kwargs.get("image_idx", 0)
        from . import utils as utilso
        frames = []          # (x_t, pred_xstart) snapshots of image 0, device-to-device copies on the stream
        x_start_rgb = x_depth = None
        pred = None
        for t in range(start_t, start_t - steps, -1):
            z = torch.randn_like(x) if t > 1 else torch.zeros_like(x)
            c_x, c_eps, c_z = self.step_coefficients(t)
            with torch.no_grad():
                pred = net(x, torch.tensor([t] * B).float().to(device))
            if record_process and (save_path is not None) and ((not t % record_every) or (t == 1)):
                # the reference's process record (:98-127): x_t and the predicted x_0 of this step.  Visualisation only,
                # a handful of steps per run - plain tensor ops on the device, no host synchronisation.
                atbar = self.alphabar[t - 1]
                x0 = float(1 / np.sqrt(atbar)) * (x[0:1] - float(np.sqrt(1 - atbar)) * pred[0:1, :image_channels])
                frames.append((x[0:1].clone(), x0))
            _lib.check(L.osm_ddpm_uncond_update(_lib.ptr(x), _lib.ptr(pred), _lib.ptr(z), c_x, c_eps, c_z, B, image_channels,
                                                pred.shape[1], H * W, _lib.stream()))
        if frames:
            xt_list = [torch.clamp(0.5 * (f[0][0, 0:3] + 1), 0, 1) for f in frames]
            rgb_list = [torch.clamp(0.5 * (f[1][0, 0:3] + 1), 0, 1) for f in frames]
            depth_list = []
            if image_channels == 4:
                for f in frames:
                    d01 = (0.5 * (f[1][0, 3] + 1)).unsqueeze(0).contiguous()
                    depth_list.append(utilso.depth_tensor_to_color_image(
                        utilso.min_max_norm_range_percentile(d01, percent_low=0.05, percent_high=0.99)))
            x_start_rgb, x_depth = rgb_list[-1], (depth_list[-1] if depth_list else None)
            from torchvision.utils import make_grid
            import torchvision.transforms.functional as tvtf
            import os
            grid = make_grid([g.cpu() for g in xt_list + rgb_list + depth_list], nrow=len(xt_list), pad_value=1.)
            tvtf.to_pil_image(grid).save(os.path.join(save_path, f"image_{image_idx}_process.png"))
        # the reference returns the last RECORDED predicted x_0 (RGB clipped to [0,1], depth percentile-normalised and
        # colour-mapped); without recording it raises UnboundLocalError (:130) - here the two entries are then None
        return x, [x_start_rgb, x_depth]
