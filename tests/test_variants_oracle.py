"""CPU: the oracle's restatement of the sampler / conditioning variants against golden vectors
produced by the unmodified reference (tests/golden/make_golden_ps.py -> ps_golden.npz): the `ps` conditioning with the
rgb_guidance operator through DDPM / DDIM `p_sample` with clip_denoised, and the `mse` loss of the osmosis conditioning.
Tolerances as in test_oracle_golden.py (same torch CPU kernels, functional vs module form: fp32 round-off)."""
import os

import numpy as np
import pytest
import torch

from oracle import osmosis_oracle as orc
from tests.golden.cases import PS_CASE, MSE_CASE, case_inputs, ps_measurement
from tests.helpers import small_state_dict, small_cfg, load_yaml_cfg, oracle_specs_from_cfg, rel_err, maxdiff

torch.set_num_threads(8)
GOLD = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ps_golden.npz")))


def ps_setup(sampler):
    cfg = load_yaml_cfg(PS_CASE["yaml"], PS_CASE["respacing"])
    d = cfg["diffusion"]
    tab = orc.make_tables(d["steps"], d["noise_schedule"], d["timestep_respacing"])
    scale = [float(v) for v in cfg["conditioning"]["params"]["scale"].split(",")]
    return cfg, tab, scale, bool(d["clip_denoised"])


@pytest.mark.parametrize("sampler", ["ddpm", "ddim"])
def test_ps_single_steps(sampler):
    cfg, tab, scale, clip = ps_setup(sampler)
    y = ps_measurement()
    for idx in PS_CASE["step_idx"]:
        x = case_inputs(f"x:ps:{idx}")
        torch.manual_seed(PS_CASE["step_seed"] + idx)
        noise = torch.randn(1, 4, *y.shape[2:])
        r = orc.ps_step(small_state_dict(), small_cfg(), tab, scale, x, y, idx, noise, sampler=sampler, clip_denoised=clip)
        pre = f"ps/{sampler}/step{idx}/"
        assert rel_err(r["loss"], GOLD[pre + "loss"]) < 1e-5
        assert maxdiff(r["pred_xstart"], GOLD[pre + "pred_xstart"]) < 1e-5
        assert maxdiff(r["x_next"], GOLD[pre + "x_next"]) < 2e-5 * max(1.0, float(np.abs(GOLD[pre + "x_next"]).max()))
        assert float(r["pred_xstart"].abs().max()) <= 1.0          # clip_denoised


@pytest.mark.parametrize("sampler", ["ddpm", "ddim"])
def test_ps_loop_matches_reference_p_sample_loop(sampler):
    """RNG order of the rgb_guidance branch: p_sample's noise [1,4,H,W] first, then the dead q_sample draw [1,3,H,W]."""
    cfg, tab, scale, clip = ps_setup(sampler)
    y = ps_measurement()
    torch.manual_seed(cfg["manual_seed"])
    x_T = torch.randn(1, 4, *y.shape[2:])

    def noise_fn(idx):
        z = torch.randn(1, 4, *y.shape[2:])
        torch.randn(1, 3, *y.shape[2:])
        return z

    x = orc.ps_sample_loop(small_state_dict(), small_cfg(), tab, scale, x_T, y, noise_fn, sampler=sampler, clip_denoised=clip)
    assert maxdiff(x, GOLD[f"ps/{sampler}/loop/img"]) < 1e-4


def test_mse_loss_single_steps():
    cfg = load_yaml_cfg(MSE_CASE["yaml"], MSE_CASE["respacing"])
    cfg["conditioning"]["params"]["loss_function"] = "mse"
    tab, op, gs, phis, names = oracle_specs_from_cfg(cfg)
    assert gs.loss_function == "mse"
    y, _ = case_inputs("meas:osmosis")
    for idx in MSE_CASE["step_idx"]:
        x = case_inputs(f"x:osmosis:{idx}")
        r = orc.guided_step(small_state_dict(), small_cfg(), tab, op, gs, x, y, phis, idx, torch.zeros_like(x))
        pre = f"mse/step{idx}/"
        assert rel_err(r["loss"], GOLD[pre + "loss"]) < 1e-4
        assert rel_err(r["grad"], GOLD[pre + "grad"]) < 2e-4
        x_t = r["mean"] - torch.tensor(gs.scale)[None, :, None, None] * torch.clamp(r["grad"], -gs.clip, gs.clip)
        assert maxdiff(x_t, GOLD[pre + "x_t"]) < 2e-5 * max(1.0, float(np.abs(GOLD[pre + "x_t"]).max()))
        for n, p in zip(names, r["phis"]):
            assert maxdiff(p, GOLD[pre + n]) < 2e-6


def test_every_mean_and_variance_processor_matches_the_reference_registries():
    """oracle.posterior(mean_type, var_type, clip_denoised) against the reference's own processor classes
    (get_mean_processor / get_var_processor) on seeded inputs, three timesteps (t = 0 gives log(0) = -inf for fixed_small)."""
    from tests.golden.cases import PROC_CASES, proc_inputs
    cfg = load_yaml_cfg(PS_CASE["yaml"], 6)
    d = cfg["diffusion"]
    tab = orc.make_tables(d["steps"], d["noise_schedule"], 6)
    x, mo = proc_inputs()
    for mean_type, var_type, clip in PROC_CASES:
        for idx in (0, 3, 5):
            x0, mean, logvar = orc.posterior(tab, idx, x, mo, clip_denoised=clip, mean_type=mean_type, var_type=var_type)
            key = f"proc/{mean_type}/{var_type}/{int(clip)}/{idx}/"
            assert maxdiff(x0, GOLD[key + "x0"]) <= 1e-6 * max(1.0, float(np.abs(GOLD[key + "x0"]).max())), key
            assert maxdiff(mean, GOLD[key + "mean"]) <= 1e-6 * max(1.0, float(np.abs(GOLD[key + "mean"]).max())), key
            lv, want = logvar.numpy(), GOLD[key + "logvar"]
            assert np.array_equal(np.isinf(lv), np.isinf(want)), key
            fin = np.isfinite(want)
            assert float(np.abs(lv[fin] - want[fin]).max(initial=0.0)) <= 1e-6 * max(1.0, float(np.abs(want[fin]).max(initial=0.0))), key


def test_adam_phi_optimizer_two_consecutive_steps():
    """`optimizer: adam` for phi (utils.py:499-500): 2 x 20 Adam steps with the state carried from one sampling step to the
    next, against the reference's torch.optim.Adam run."""
    cfg = load_yaml_cfg(MSE_CASE["yaml"], MSE_CASE["respacing"])
    cfg["measurement"]["operator"]["optimizer"] = "adam"
    tab, op, gs, phis, names = oracle_specs_from_cfg(cfg)
    assert op.optimizer == "adam"
    y, _ = case_inputs("meas:osmosis")
    state = None
    for idx in (2, 1):
        x = case_inputs(f"x:osmosis:{idx}")
        r = orc.guided_step(small_state_dict(), small_cfg(), tab, op, gs, x, y, phis, idx, torch.zeros_like(x), opt_state=state)
        phis, state = r["phis"], r["opt_state"]
        pre = f"adam/step{idx}/"
        assert rel_err(r["loss"], GOLD[pre + "loss"]) < 1e-4
        for n, p in zip(names, phis):
            assert maxdiff(p, GOLD[pre + n]) < 2e-6, (idx, n)
        x_t = r["mean"] - torch.tensor(gs.scale)[None, :, None, None] * torch.clamp(r["grad"], -gs.clip, gs.clip)
        assert maxdiff(x_t, GOLD[pre + "x_t"]) < 2e-5 * max(1.0, float(np.abs(GOLD[pre + "x_t"]).max()))
    assert state[0]["step"] == 40
