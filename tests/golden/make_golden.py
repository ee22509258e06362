"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference imports matplotlib and natsort at module top level for visualisation; neither is
installed here, so empty stand-in modules are registered before the import (SURVEY.md hazard 4).
Nothing under /root/reference is modified or copied.  Only OUTPUT tensors are stored; inputs and the
synthetic weights are regenerated from seeds by the tests (osmosis_diffusion_code_b200/synthetic.py).
"""
import argparse
import json
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
from guided_diffusion.nn import timestep_embedding  # noqa: E402
import yaml  # noqa: E402

from osmosis_diffusion_code_b200.synthetic import synth_state_dict, synth_measurement  # noqa: E402
from tests.golden.cases import SMALL_UNET, CASES, case_inputs  # noqa: E402

torch.set_num_threads(8)


def ref_model():
    m = create_model(**SMALL_UNET, model_path="/nonexistent")
    specs = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    sd = synth_state_dict(specs, SMALL_UNET["num_channels"], seed=7, delta=0.05)
    m.load_state_dict(sd, strict=True)
    return m.eval(), specs


def main():
    out = {}
    model, specs = ref_model()
    with open(os.path.join(HERE, "small_unet_param_specs.json"), "w") as f:
        json.dump([[k, list(s)] for k, s in specs], f)

    # ---- G0: timestep embedding ----
    t = torch.tensor([0, 1, 37, 512, 999])
    out["temb"] = timestep_embedding(t, 256).numpy()
    out["temb_float"] = timestep_embedding(torch.tensor([1.0, 500.0, 1000.0]), 256).numpy()

    # ---- G1: UNet forward + input gradient ----
    x, tt, cot = case_inputs("unet")
    xg = x.clone().requires_grad_(True)
    y = model(xg, tt)
    (gx,) = torch.autograd.grad(y, xg, cot)
    out["unet_out"] = y.detach().numpy()
    out["unet_gx"] = gx.numpy()

    # ---- per-config golden: posterior, operator, one frozen + one optimised guided step, short loop ----
    for cname, c in CASES.items():
        cfg = yaml.load(open(os.path.join(ROOT, "configs", c["yaml"])), Loader=yaml.FullLoader)
        cfg["diffusion"]["timestep_respacing"] = c["respacing"]
        B = 1
        opcfg = dict(cfg["measurement"]["operator"]); opcfg["batch_size"] = B
        sampler = create_sampler(**cfg["diffusion"])
        y_meas, x_gt = case_inputs("meas:" + cname)

        def fresh():
            op = get_operator(device="cpu", **opcfg)
            noiser = get_noise(**cfg["measurement"]["noise"])
            cond = get_conditioning_method(cfg["conditioning"]["method"], op, noiser, **cfg["conditioning"]["params"],
                                           **cfg["sample_pattern"], **cfg["aux_loss"])
            return op, cond

        op, cond = fresh()
        out[f"{cname}/op_fwd"] = op.forward(x_gt).detach().numpy()

        # posterior at two respaced indices
        T = sampler.num_timesteps
        for idx in c["post_idx"]:
            xi = case_inputs(f"x:{cname}:{idx}")
            with torch.no_grad():
                o = sampler.p_mean_variance(model, xi, torch.tensor([idx]))
            for k in ("mean", "log_variance", "pred_xstart"):
                out[f"{cname}/post{idx}/{k}"] = o[k].numpy()

        # single guided steps (frozen / optimised phase) through the reference's own calls
        for idx in c["step_idx"]:
            op, cond = fresh()
            img = case_inputs(f"x:{cname}:{idx}").clone().requires_grad_(True)
            time = torch.tensor([idx])
            o = sampler.p_mean_variance(model, img, time)
            from osmosis_utils.utils import is_freeze_phi
            freeze = is_freeze_phi(cfg["sample_pattern"], idx, T)
            x_t, loss, vd, grads, aux = cond.conditioning(x_t=o["mean"], measurement=y_meas, noisy_measurement=None,
                                                          x_prev=img, x_0_hat=o["pred_xstart"], freeze_phi=freeze,
                                                          time_index=float(idx) / T)
            noise = case_inputs(f"noise:{cname}:{idx}")
            x_next = x_t.detach().clone()
            if idx != 0:
                x_next += torch.exp(0.5 * o["log_variance"].detach()) * noise
            out[f"{cname}/step{idx}/x_next"] = x_next.numpy()
            out[f"{cname}/step{idx}/grad"] = grads.numpy()
            out[f"{cname}/step{idx}/loss"] = np.asarray(loss, dtype=np.float32)
            out[f"{cname}/step{idx}/freeze"] = np.asarray([int(freeze)])
            for k, v in vd.items():
                out[f"{cname}/step{idx}/{k}"] = v.detach().numpy()

        # the reference's own p_sample_loop, RNG from manual_seed (order: SURVEY Appendix C)
        op, cond = fresh()
        torch.manual_seed(cfg["manual_seed"])
        x_start = torch.randn(1, 4, *y_meas.shape[2:]).requires_grad_()
        img, vd, loss, x0 = sampler.p_sample_loop(model=model, x_start=x_start, measurement=y_meas,
                                                  measurement_cond_fn=cond.conditioning, record=False, save_root=None,
                                                  pretrain_model="osmosis", rgb_guidance=False,
                                                  sample_pattern=cfg["sample_pattern"])
        out[f"{cname}/loop/img"] = img.detach().numpy()
        out[f"{cname}/loop/pred_xstart"] = x0.numpy()
        out[f"{cname}/loop/loss"] = np.asarray(loss, dtype=np.float32)
        for k, v in vd.items():
            out[f"{cname}/loop/{k}"] = v.detach().numpy()
        print(cname, "done", {k: v.flatten().tolist() for k, v in vd.items()}, loss)

    np.savez_compressed(os.path.join(HERE, "small_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
