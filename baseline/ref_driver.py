"""Drives the UNMODIFIED reference (osmosis-diffusion/osmosis-diffusion-code) for the benchmark's reference arms and for
the like-for-like GPU parity tests.

The reference is a plain Python tree; `install()` copies it verbatim to `baseline/_ref/` (git-ignored, shipped to the GPU
box by gpurun) - there is nothing to compile.  It imports `matplotlib` and `natsort` at module top level for
visualisation only; neither is installed, so empty stand-in modules are registered before the import.  Nothing of this
repo's package, kernels or oracle is on the paths below: the model, sampler, operator and conditioning objects are the
reference's own classes, called through its own public API in the sequence `osmosis_sampling.py:65-67, 145-164, 194-204`
uses.  Only the INPUTS are shared with the native arm: synthetic denoiser-like weights and a synthetic scene
(`osmosis_diffusion_code_b200/synthetic.py`, pure CPU torch).
"""
from __future__ import annotations

import contextlib
import os
import shutil
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"
_MARK = os.path.join(REF, "guided_diffusion", "gaussian_diffusion.py")


def install(verbose=False):
    """Copy the reference tree (without its README figures) to baseline/_ref.  Only possible where /root/reference exists
    (the build container); on the GPU box the copy that travelled with the snapshot is used."""
    if os.path.exists(_MARK):
        return True
    if not os.path.isdir(REF_SRC):
        return False
    shutil.copytree(REF_SRC, REF, ignore=shutil.ignore_patterns("figures", ".git", "__pycache__"), dirs_exist_ok=True)
    if verbose:
        print("installed the reference tree at", REF)
    return True


def available():
    return os.path.exists(_MARK)


_mods = None


def import_reference():
    """Returns a namespace of the reference's own modules (imported from baseline/_ref)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python -c 'from baseline import ref_driver; ref_driver.install()'` "
                           "in the build container")
    for name in ("matplotlib", "matplotlib.pyplot", "natsort"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["natsort"].natsorted = sorted
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import guided_diffusion.unet as r_unet
    import guided_diffusion.gaussian_diffusion as r_gd
    import guided_diffusion.measurements as r_meas
    import guided_diffusion.condition_methods as r_cond
    import osmosis_utils.utils as r_utils
    assert os.path.realpath(r_gd.__file__).startswith(os.path.realpath(REF)), r_gd.__file__
    _mods = types.SimpleNamespace(unet=r_unet, gd=r_gd, meas=r_meas, cond=r_cond, utils=r_utils)
    return _mods


def reference_model(args, device, seed=7, delta=0.05):
    """The reference's `create_model(**cfg.unet_model)` with the synthetic weights loaded through its own load_state_dict."""
    import torch
    from osmosis_diffusion_code_b200.synthetic import synth_state_dict
    R = import_reference()
    um = dict(args.unet_model)
    um["model_path"] = "/nonexistent/osmosis_outdoor.pt"
    with contextlib.redirect_stdout(sys.stderr):   # the reference prints the failed checkpoint load and keeps the random init
        model = R.unet.create_model(**um)
    specs = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth_state_dict(specs, um["num_channels"], seed=seed, delta=delta), strict=True)
    return model.to(device).eval()


def reference_pieces(args, device, batch=1):
    """operator / noiser / conditioning method / sampler exactly as osmosis_sampling.py:145-164 builds them per image."""
    R = import_reference()
    opc = dict(args.measurement["operator"])
    opc["batch_size"] = batch
    operator = R.meas.get_operator(device=device, **opc)
    noiser = R.meas.get_noise(**args.measurement["noise"])
    cond = R.cond.get_conditioning_method(args.conditioning["method"], operator, noiser, **args.conditioning["params"],
                                          **args.sample_pattern, **args.aux_loss)
    sampler = R.gd.create_sampler(**args.diffusion)
    return operator, cond, sampler


def reference_measurement(args, device, index=0, size=256):
    """y = 2 A_phi0(x_gt) - 1 with the reference's own operator at the YAML's initial phi (SURVEY 8(d) recipe)."""
    import torch
    from osmosis_diffusion_code_b200.synthetic import synth_scene
    R = import_reference()
    opc = dict(args.measurement["operator"])
    opc["batch_size"] = 1
    op = R.meas.get_operator(device="cpu", **opc)
    with torch.no_grad():
        y = 2 * op.forward(synth_scene(index, size)) - 1
    return y.float().to(device)


class _Stop(Exception):
    pass


class _TimedModel:
    """Stamps the host clock (after a device sync on CUDA) at every UNet call = the start of every reverse step."""

    def __init__(self, model, device, budget_s, min_steps):
        self.model, self.cuda = model, str(device).startswith("cuda")
        self.stamps, self.budget_s, self.min_steps, self.t_begin = [], budget_s, min_steps, time.perf_counter()

    def __call__(self, *a, **k):
        if self.cuda:
            import torch
            torch.cuda.synchronize()
        now = time.perf_counter()
        if self.budget_s is not None and now - self.t_begin > self.budget_s and len(self.stamps) >= self.min_steps:
            raise _Stop()
        self.stamps.append(now)
        return self.model(*a, **k)


def time_reference_chain(config_path, device, steps, warmup, size=256, budget_s=None, threads=None):
    """Runs the reference's own `p_sample_loop` on a (steps + warmup)-step respacing of the config's chain (the same
    chain the native arm times) and returns per-step wall times of the `steps` steps after the first `warmup`.
    `budget_s` stops the loop early (after at least warmup + 2 steps) by raising out of the model call."""
    import torch
    R = import_reference()
    if threads:
        torch.set_num_threads(threads)
    args = R.utils.arguments_from_file(config_path)
    base_T = int(args.diffusion["steps"])
    args.diffusion = dict(args.diffusion)
    args.diffusion["timestep_respacing"] = min(steps + warmup, base_T)
    model = reference_model(args, device)
    operator, cond, sampler = reference_pieces(args, device, batch=1)
    y = reference_measurement(args, device, 0, size)
    if getattr(args, "degamma_input", False):
        y = 2 * torch.pow(0.5 * (y + 1), 2.2) - 1
    timed = _TimedModel(model, device, budget_s, warmup + 2)
    torch.manual_seed(args.manual_seed)
    x_start = torch.randn([1, 4, size, size], device=device).requires_grad_()
    done = None
    try:
        done = sampler.p_sample_loop(model=timed, x_start=x_start, measurement=y, measurement_cond_fn=cond.conditioning,
                                     pretrain_model=args.unet_model["pretrain_model"], rgb_guidance=args.rgb_guidance,
                                     sample_pattern=args.sample_pattern, record=False, save_root=None, image_idx=0,
                                     record_every=args.record_every, original_file_name="bench", save_grids_path=None,
                                     global_iteration=0)
    except _Stop:
        pass
    if timed.cuda:
        torch.cuda.synchronize()
    end = time.perf_counter()
    stamps = timed.stamps + ([end] if done is not None else [])
    per_step = [b - a for a, b in zip(stamps[:-1], stamps[1:])]
    finite = bool(torch.isfinite(done[0]).all()) if done is not None else None
    return dict(step_s=per_step[warmup:], warmup_s=per_step[:warmup], T=sampler.num_timesteps, finite=finite,
                completed=done is not None)
