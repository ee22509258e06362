"""Golden vectors of the sampler / conditioning variants, produced by the UNMODIFIED reference:

  * the `ps` conditioning (condition_methods.py:234-251) with the `rgb_guidance` operator (measurements.py:80-97) through
    `DDPM.p_sample` / `DDIM.p_sample` (gaussian_diffusion.py:492-535) with `clip_denoised: True`
    (posterior_mean_variance.py:41-50) - rgb_guidance_sample_config.yaml: single steps and the 6-step `p_sample_loop`
    with `rgb_guidance=True` (gaussian_diffusion.py:232-233, 299-306);
  * `loss_function: mse` of the osmosis conditioning (condition_methods.py:133-138): single steps.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_ps.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for name in ("matplotlib", "matplotlib.pyplot", "natsort"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["natsort"].natsorted = sorted
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")

from guided_diffusion.unet import create_model  # noqa: E402  (the reference)
from guided_diffusion.gaussian_diffusion import create_sampler  # noqa: E402
from guided_diffusion.measurements import get_operator, get_noise  # noqa: E402
from guided_diffusion.condition_methods import get_conditioning_method  # noqa: E402
from osmosis_utils.utils import is_freeze_phi  # noqa: E402
import yaml  # noqa: E402

from osmosis_diffusion_code_b200.synthetic import synth_state_dict  # noqa: E402
from tests.golden.cases import SMALL_UNET, PS_CASE, MSE_CASE, case_inputs, ps_measurement  # noqa: E402

torch.set_num_threads(8)


def ref_model():
    m = create_model(**SMALL_UNET, model_path="/nonexistent")
    specs = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    m.load_state_dict(synth_state_dict(specs, SMALL_UNET["num_channels"], seed=7, delta=0.05), strict=True)
    return m.eval()


def main():
    out = {}
    model = ref_model()
    # ---------------------------------------------------------------- ps / rgb_guidance
    cfg = yaml.load(open(os.path.join(ROOT, "configs", PS_CASE["yaml"])), Loader=yaml.FullLoader)
    y = ps_measurement()
    for sname in ("ddpm", "ddim"):
        d = dict(cfg["diffusion"]); d["timestep_respacing"] = PS_CASE["respacing"]; d["sampler"] = sname
        sampler = create_sampler(**d)
        op = get_operator(device="cpu", **cfg["measurement"]["operator"])
        noiser = get_noise(**cfg["measurement"]["noise"])
        cond = get_conditioning_method(cfg["conditioning"]["method"], op, noiser, **cfg["conditioning"]["params"])
        for idx in PS_CASE["step_idx"]:
            img = case_inputs(f"x:ps:{idx}").clone().requires_grad_(True)
            torch.manual_seed(PS_CASE["step_seed"] + idx)              # p_sample draws randn_like(x) itself
            o = sampler.p_sample(x=img, t=torch.tensor([idx]), model=model)
            x_next, loss = cond.conditioning(x_t=o["sample"], measurement=y, noisy_measurement=None, x_prev=img,
                                             x_0_hat=o["pred_xstart"])
            out[f"ps/{sname}/step{idx}/x_next"] = x_next.detach().numpy()
            out[f"ps/{sname}/step{idx}/pred_xstart"] = o["pred_xstart"].detach().numpy()
            out[f"ps/{sname}/step{idx}/loss"] = np.asarray([float(loss)], dtype=np.float32)
        torch.manual_seed(cfg["manual_seed"])
        x_start = torch.randn(1, 4, *y.shape[2:]).requires_grad_()
        img = sampler.p_sample_loop(model=model, x_start=x_start, measurement=y, measurement_cond_fn=cond.conditioning,
                                    record=False, save_root=None, pretrain_model="osmosis", rgb_guidance=True,
                                    sample_pattern=cfg["sample_pattern"])
        out[f"ps/{sname}/loop/img"] = img.detach().numpy()
        print("ps", sname, "done", float(img.abs().max()))
    # ---------------------------------------------------------------- osmosis conditioning with the mse loss
    cfg = yaml.load(open(os.path.join(ROOT, "configs", MSE_CASE["yaml"])), Loader=yaml.FullLoader)
    cfg["diffusion"]["timestep_respacing"] = MSE_CASE["respacing"]
    cfg["conditioning"]["params"]["loss_function"] = "mse"
    sampler = create_sampler(**cfg["diffusion"])
    T = sampler.num_timesteps
    y_meas, _ = case_inputs("meas:osmosis")
    for idx in MSE_CASE["step_idx"]:
        opcfg = dict(cfg["measurement"]["operator"]); opcfg["batch_size"] = 1
        op = get_operator(device="cpu", **opcfg)
        cond = get_conditioning_method(cfg["conditioning"]["method"], op, get_noise(**cfg["measurement"]["noise"]),
                                       **cfg["conditioning"]["params"], **cfg["sample_pattern"], **cfg["aux_loss"])
        img = case_inputs(f"x:osmosis:{idx}").clone().requires_grad_(True)
        o = sampler.p_mean_variance(model, img, torch.tensor([idx]))
        freeze = is_freeze_phi(cfg["sample_pattern"], idx, T)
        x_t, loss, vd, grads, aux = cond.conditioning(x_t=o["mean"], measurement=y_meas, noisy_measurement=None, x_prev=img,
                                                      x_0_hat=o["pred_xstart"], freeze_phi=freeze, time_index=float(idx) / T)
        out[f"mse/step{idx}/x_t"] = x_t.detach().numpy()
        out[f"mse/step{idx}/grad"] = grads.numpy()
        out[f"mse/step{idx}/loss"] = np.asarray(loss, dtype=np.float32)
        for k, v in vd.items():
            out[f"mse/step{idx}/{k}"] = v.detach().numpy()
        print("mse", idx, "freeze", freeze, loss)
    # ---------------------------------------------------------------- Adam for phi (utils.py:499-500): two consecutive
    # optimised steps with ONE operator, so the optimizer state carries over as in a sampling run
    cfg = yaml.load(open(os.path.join(ROOT, "configs", MSE_CASE["yaml"])), Loader=yaml.FullLoader)
    cfg["diffusion"]["timestep_respacing"] = MSE_CASE["respacing"]
    sampler = create_sampler(**cfg["diffusion"])
    opcfg = dict(cfg["measurement"]["operator"]); opcfg["batch_size"] = 1; opcfg["optimizer"] = "adam"
    op = get_operator(device="cpu", **opcfg)
    cond = get_conditioning_method(cfg["conditioning"]["method"], op, get_noise(**cfg["measurement"]["noise"]),
                                   **cfg["conditioning"]["params"], **cfg["sample_pattern"], **cfg["aux_loss"])
    for idx in (2, 1):
        img = case_inputs(f"x:osmosis:{idx}").clone().requires_grad_(True)
        o = sampler.p_mean_variance(model, img, torch.tensor([idx]))
        x_t, loss, vd, grads, aux = cond.conditioning(x_t=o["mean"], measurement=y_meas, noisy_measurement=None, x_prev=img,
                                                      x_0_hat=o["pred_xstart"], freeze_phi=False, time_index=float(idx) / T)
        out[f"adam/step{idx}/x_t"] = x_t.detach().numpy()
        out[f"adam/step{idx}/loss"] = np.asarray(loss, dtype=np.float32)
        for k, v in vd.items():
            out[f"adam/step{idx}/{k}"] = v.detach().numpy().copy()
        print("adam", idx, loss, {k: v.flatten().tolist() for k, v in vd.items()})
    # ---------------------------------------------------------------- every mean / variance processor of the registries
    from guided_diffusion.posterior_mean_variance import get_mean_processor, get_var_processor
    from tests.golden.cases import PROC_CASES, proc_inputs
    d = dict(cfg["diffusion"]); d["timestep_respacing"] = 6
    betas = create_sampler(**d).betas
    x, mo = proc_inputs()
    for mean_type, var_type, clip in PROC_CASES:
        mp = get_mean_processor(mean_type, betas=betas, dynamic_threshold=False, clip_denoised=clip)
        vp = get_var_processor(var_type, betas=betas)
        for idx in (0, 3, 5):
            t = torch.tensor([idx] * x.shape[0])
            mean, x0 = mp.get_mean_and_xstart(x, t, mo[:, :4])
            var, logvar = vp.get_variance(mo[:, 4:], t)
            key = f"proc/{mean_type}/{var_type}/{int(clip)}/{idx}/"
            out[key + "mean"], out[key + "x0"] = mean.numpy(), x0.numpy()
            out[key + "logvar"] = np.broadcast_to(logvar.numpy(), x.shape).astype(np.float32).copy()
    np.savez_compressed(os.path.join(HERE, "ps_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
