"""Synthetic inputs for benchmarks and parity tests: denoiser-like UNet weights and physically
consistent measurements.

There is no network access and the upstream checkpoint (`osmosis_outdoor.pt`) is a download, so every
benchmark runs on random-init weights "of that architecture".  Plain random init makes the chain
non-finite after one step (|x0_hat| ~ 700 at t=999 -> exp overflow in the operator; SURVEY.md hazard 2)
and leaves 131 of 356 tensors exactly zero (hazard 3), so the weights are drawn so that
    eps(x, t) = x / rms(x) + delta * R(x, t)
with R the fully random network (SURVEY.md Appendix F describes the construction).  The state_dict has
the guided-diffusion key names / OIHW shapes, so the same dict loads into the upstream `UNetModel`,
into the CPU oracle and into the native engine.

Everything is generated on the CPU with a seeded `torch.Generator` so that the container that made the
golden fixtures and the GPU box regenerate bit-identical tensors.
"""
from __future__ import annotations

import math

import torch


def synth_state_dict(param_specs, model_channels: int, in_channels: int = 4, seed: int = 7, delta: float = 0.05):
    """param_specs: iterable of (name, shape) in state_dict order.  Returns {name: fp32 CPU tensor}."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_specs:
        shape = tuple(shape)
        leaf = name.rsplit(".", 1)[-1]
        is_norm = len(shape) == 1 and (".in_layers.0." in name or ".out_layers.0." in name or ".norm." in name
                                       or name.startswith("out.0."))
        small = ".out_layers.3." in name or ".proj_out." in name or name.startswith("out.2.")
        if is_norm:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
        elif small:  # tensors upstream zero-initialises: give them a small random value so they matter
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            if leaf == "weight":
                fan_in = math.prod(shape[1:])
            else:  # bias: fan_in of the matching weight is not known here; 1/sqrt(C_out) keeps it small
                fan_in = shape[0]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.float().contiguous()

    # ---- denoiser surgery: eps_c = x_c / rms + delta * random features ----
    C = model_channels
    cpg = C // 32
    if cpg >= 8:
        P, N, dead, gain = [0, 1, 2, 3], [4, 5, 6, 7], list(range(8, cpg)), math.sqrt(8.0 / cpg)
    elif cpg in (2, 4):
        P, N, dead, gain = [0, 2, 4, 6], [1, 3, 5, 7], [], 1.0
    else:
        raise ValueError("model_channels must give 2, 4 or >= 8 channels per GroupNorm group")
    w_in, b_in = sd["input_blocks.0.0.weight"], sd["input_blocks.0.0.bias"]
    for c in range(in_channels):
        for idx, sgn in ((P[c], 1.0), (N[c], -1.0)):
            w_in[idx].zero_(); w_in[idx, c, 1, 1] = sgn; b_in[idx] = 0
    last = max(int(k.split(".")[1]) for k in sd if k.startswith("output_blocks."))
    p = f"output_blocks.{last}.0"
    sk_w, sk_b = sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"]
    oc_w, oc_b = sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"]
    off = sk_w.shape[1] - C  # the hs[0] half of the last concat starts here
    for idx in P + N + dead:
        sk_w[idx].zero_(); sk_b[idx] = 0; oc_w[idx].zero_(); oc_b[idx] = 0
    for idx in P + N:
        sk_w[idx, off + idx, 0, 0] = 1.0
    gn_w, gn_b = sd["out.0.weight"], sd["out.0.bias"]
    fc_w, fc_b = sd["out.2.weight"], sd["out.2.bias"]
    sel = P + N + dead
    gn_w[sel] = 1; gn_b[sel] = 0
    fc_w.mul_(delta); fc_b.mul_(delta)  # out.2 was drawn at std 0.02; scaled by delta like Appendix F
    fc_w[:, sel] = 0
    for c in range(in_channels):
        fc_w[c, P[c], 1, 1] = gain; fc_w[c, N[c], 1, 1] = -gain
    return sd


def synth_scene(index: int, size: int = 256):
    """Smooth random scene: returns x_gt [1,4,size,size] in [-1,1] (RGB + depth), generated on the CPU."""
    g = torch.Generator().manual_seed(1000 + index)
    J = torch.nn.functional.interpolate(torch.rand(1, 3, 16, 16, generator=g), size=size, mode="bilinear", align_corners=False)
    d = torch.nn.functional.interpolate(torch.rand(1, 1, 8, 8, generator=g), size=size, mode="bilinear", align_corners=False)
    return torch.cat([2 * J - 1, 2 * d - 1], dim=1)


def synth_measurement(index: int, size: int, phi_a, phi_b, phi_inf, depth_type="gamma", value=(1.4, 1.4, 1.0)):
    """y = 2 * A_phi(x_gt) - 1 for scene `index` (the underwater image-formation model, CPU fp32).

    phi_* are length-3 sequences (or a scalar phi_a == phi_b for the haze model).  Returns (y [1,3,H,W], x_gt).
    """
    x = synth_scene(index, size)
    J = 0.5 * (x[:, :3] + 1)
    d = x[:, 3:4]
    if depth_type == "gamma":
        d = torch.pow((d + value[0]) * value[1], value[2])
    elif depth_type in (None, "original"):
        d = 0.5 * (d + 1.0)
    else:
        d = d + value
    as_t = lambda p: torch.as_tensor(p, dtype=torch.float32).reshape(1, -1, 1, 1)
    pa, pb, pinf = as_t(phi_a), as_t(phi_b), as_t(phi_inf)
    uw = J * torch.exp(-pa * d) + pinf * (1 - torch.exp(-pb * d))
    return 2 * uw - 1, x
