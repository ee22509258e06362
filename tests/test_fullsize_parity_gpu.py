"""GPU, BASELINE size (the shipped 553 M-parameter UNet at 256x256) against the CPU oracle and against the UNMODIFIED
reference running on the same GPU (baseline/_ref, eager CUDA).

What only exists at this size - 1024/1536/2048-channel convs through the split-K cluster plan, the CTA-pair conv kernel,
flash attention at L = 1024 (8 heads) / 256 / 64 (16 heads), the six-level skip-concat aliasing - is pinned here, not only
through the small golden UNet.

Tolerances (normalised max error = max|a-b| / max|b| unless stated), per level:
  * UNet output / input-VJP vs the fp32 CPU oracle:        exact mode 2e-4 / 1e-3,  product mode (TF32 tensor cores) 1e-2 / 3e-2
  * guided step (teacher-forced: the oracle's x_t, phi, noise), vs the oracle and vs the reference on CUDA:
      pred_xstart 1e-4 exact / 1e-2 product TIMES the step's amplification sqrt(1/abar_t - 1) max|x_t| / max|x0| (x0 = sqrt(1/abar) x -
      sqrt(1/abar - 1) eps magnifies the UNet's rounding: 38x at t = 850), loss 1e-3 / 2e-2,
      x_{t-1}: every pixel within the clamp bound 2 * scale * clip, >= 99 % (exact) / 97 % (product) of pixels within 1e-3 / 2e-2,
      phi 5e-6 exact / 5e-5 product (absolute)
  * RNG stream: the loop's draws are BIT-identical to the reference's (`randn_like` of [1,3,H,W] then [1,4,H,W] per step on the device).
"""
import contextlib
import os
import sys

import pytest
import torch

from oracle import osmosis_oracle as orc
from baseline import ref_driver as rd
from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file, is_freeze_phi, load_yaml
from osmosis_diffusion_code_b200.guided_diffusion.unet import create_model
from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import create_sampler, FusedStepper
from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator, get_noise
from osmosis_diffusion_code_b200.guided_diffusion.condition_methods import get_conditioning_method
from osmosis_diffusion_code_b200.synthetic import synth_state_dict, synth_measurement, synth_scene
from tests.helpers import rel_err, maxdiff

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV, S = "cuda", 256
CONFIGS = ["osmosis_sample_config.yaml", "osmosis_simulation_sample_config.yaml", "osmosis_haze_sample_config.yaml"]
needs_ref = pytest.mark.skipif(not rd.available(), reason="baseline/_ref (copy of the unmodified reference) is not present")
_cache = {}


def cfg_path(name):
    return os.path.join(ROOT, "configs", name)


def state_dict():
    if "sd" not in _cache:
        a = arguments_from_file(cfg_path(CONFIGS[0]))
        ucfg = orc.UNetConfig.from_create_model_kwargs(**a.unet_model)
        _cache["ucfg"] = ucfg
        _cache["sd"] = synth_state_dict(list(orc.param_shapes(ucfg).items()), ucfg.model_channels, seed=7, delta=0.05)
    return _cache["sd"], _cache["ucfg"]


def full_model(conv_mode):
    if ("m", conv_mode) not in _cache:
        a = arguments_from_file(cfg_path(CONFIGS[0]))
        um = dict(a.unet_model); um["model_path"] = ""
        with contextlib.redirect_stdout(sys.stderr):
            m = create_model(**um, conv_mode=conv_mode)
        m.load_state_dict(state_dict()[0])
        _cache[("m", conv_mode)] = m.to(DEV).eval()
    return _cache[("m", conv_mode)]


def ref_model():
    if "ref" not in _cache:
        R = rd.import_reference()
        args = R.utils.arguments_from_file(cfg_path(CONFIGS[0]))
        _cache["ref"] = rd.reference_model(args, DEV)
    return _cache["ref"]


def native_objects(cfg_name, B, respacing=1000):
    a = arguments_from_file(cfg_path(cfg_name))
    opc = dict(a.measurement["operator"]); opc["batch_size"] = B
    op = get_operator(device=DEV, **opc)
    cond = get_conditioning_method(a.conditioning["method"], op, get_noise(**a.measurement["noise"]), **a.conditioning["params"],
                                   **a.sample_pattern, **a.aux_loss)
    d = dict(a.diffusion); d["timestep_respacing"] = respacing
    return a, op, cond, create_sampler(**d)


def measurement(a, index=0):
    opc = a.measurement["operator"]
    ph = lambda k, dflt: [float(v) for v in str(opc.get(k, dflt)).split(",")]
    pa, pb = (ph("phi_a", "1"), ph("phi_b", "1")) if "phi_a" in opc else (ph("phi_ab", "1"), ph("phi_ab", "1"))
    return synth_measurement(index, S, pa, pb, ph("phi_inf", "0.2,0.4,0.7"), depth_type=opc.get("depth_type"))[0]


def chain_state(tab, idx, seed):
    """A plausible x_t of the chain: sqrt(abar_t) x_gt + sqrt(1 - abar_t) eps, plus the step's noise (CPU, seeded)."""
    g = torch.Generator().manual_seed(seed)
    abar = float(tab.alphas_cumprod[idx])
    x = abar ** 0.5 * synth_scene(3, S) + (1 - abar) ** 0.5 * torch.randn(1, 4, S, S, generator=g)
    return x.float(), torch.randn(1, 4, S, S, generator=g)


# ------------------------------------------------------------------------------------------------- UNet vs the CPU oracle


def test_unet_forward_and_input_vjp_vs_oracle_at_full_size():
    sd, ucfg = state_dict()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 4, S, S, generator=g)
    t = torch.tensor([500])
    cot = torch.randn(1, 8, S, S, generator=g) * 1e-3
    xg = x.clone().requires_grad_(True)
    want = orc.unet_forward(sd, ucfg, xg, t)
    (gwant,) = torch.autograd.grad(want, xg, cot)
    want = want.detach()
    for mode, tol_o, tol_g in (("fp32", 2e-4, 1e-3), ("tc", 1e-2, 3e-2)):
        m = full_model(mode)
        out = m._forward_raw(x.to(DEV), t.to(DEV).float()).clone()
        gx = m._vjp_raw(cot.to(DEV)).clone()
        torch.cuda.synchronize()
        eo, eg = rel_err(out.cpu(), want), rel_err(gx.cpu(), gwant)
        print(f"full-size UNet vs oracle [{mode}]: out {eo:.2e} (tol {tol_o}), input-VJP {eg:.2e} (tol {tol_g})")
        assert eo < tol_o and eg < tol_g, (mode, eo, eg)


# ------------------------------------------------------------------------------------------------- guided steps


def _native_step(mode, cfg_name, idx, x, noise, y):
    a, op, cond, sampler = native_objects(cfg_name, 1)
    m = full_model(mode)
    img = x.to(DEV).clone()
    st = sampler.fused_state(m, cond, img, y.to(DEV))
    freeze = is_freeze_phi(a.sample_pattern, idx, sampler.num_timesteps)
    st["t_idx"].fill_(idx); st["t_model"].fill_(sampler._model_timestep(idx)); st["freeze"].fill_(int(freeze))
    sampler.fused_step(m, cond, st, img, noise.to(DEV))
    torch.cuda.synchronize()
    return dict(x_next=img.cpu(), x0=st["x0"].cpu(), loss=st["losses"][:, 0].cpu(), grad=st["grad"].cpu(), logvar=st["logvar"].cpu(),
                phi={n: getattr(op, n).detach().cpu().reshape(-1) for n in op.groups}, cond=cond, freeze=freeze)


def _amp(tab, idx, x, want_x0):
    return max(1.0, float(tab.sqrt_recipm1_alphas_cumprod[idx]) * float(x.abs().max()) / float(want_x0.abs().max()))


def _check_step(tag, got, want_x_next, want_x0, want_loss, want_phi, exact, cond, amp):
    bound = 2 * float(cond.scale.max()) * cond.gradient_clip_value * 1.01
    scale = max(1.0, float(want_x_next.abs().max()))
    d = (got["x_next"] - want_x_next).abs()
    tight = (1e-3 if exact else 2e-2) * scale
    frac = float((d < tight).float().mean())
    e_x0, e_loss = rel_err(got["x0"], want_x0), rel_err(got["loss"], want_loss)
    e_phi = max(maxdiff(got["phi"][n], want_phi[n].reshape(-1)) for n in got["phi"])
    print(f"{tag}: x0 {e_x0:.2e} (amplification {amp:.1f})  loss {e_loss:.2e}  x_next max {float(d.max()):.2e} (clamp bound {bound:.3f}), within {tight:.0e}: {frac:.4f}  "
          f"phi {e_phi:.2e}")
    assert e_x0 < (1e-4 if exact else 1e-2) * amp, tag
    assert e_loss < (1e-3 if exact else 2e-2), tag
    assert float(d.max()) <= bound + tight, tag
    assert frac > (0.99 if exact else 0.97), tag
    assert e_phi < (5e-6 if exact else 5e-5), tag


@pytest.mark.parametrize("cfg_name", CONFIGS)
def test_teacher_forced_guided_steps_vs_oracle_at_full_size(cfg_name):
    sd, ucfg = state_dict()
    cfg = load_yaml(cfg_path(cfg_name))
    tab, ospec, gspec, phis, names = orc.specs_from_config(cfg, 1)
    a = arguments_from_file(cfg_path(cfg_name))
    y = measurement(a)
    for idx in (850, 400):                      # a frozen-phi step and an optimised-phi step (20 inner iterations)
        x, noise = chain_state(tab, idx, seed=idx)
        r = orc.guided_step(sd, ucfg, tab, ospec, gspec, x, y, phis, idx, noise)
        want_phi = {n: p for n, p in zip(names, r["phis"])}
        for mode in ("fp32", "tc"):
            got = _native_step(mode, cfg_name, idx, x, noise, y)
            assert got["freeze"] == orc.is_freeze_phi(gspec, idx, tab.num_timesteps)
            _check_step(f"{cfg_name} t={idx} [{mode}] vs oracle", got, r["x_next"], r["pred_xstart"], r["loss"], want_phi, mode == "fp32",
                        got["cond"], _amp(tab, idx, x, r["pred_xstart"]))


@needs_ref
@pytest.mark.parametrize("cfg_name", CONFIGS)
def test_teacher_forced_guided_steps_vs_reference_on_cuda(cfg_name):
    """The same steps against the unmodified reference executing on this GPU: `p_mean_variance` + `conditioning`
    (gaussian_diffusion.py:345-365, condition_methods.py:146-231), once with cuDNN's default TF32 convolutions (what a
    reference user gets; compared with the product mode) and once with TF32 switched off (compared with the exact mode)."""
    R = rd.import_reference()
    args = R.utils.arguments_from_file(cfg_path(cfg_name))
    model = ref_model()
    tab = orc.specs_from_config(load_yaml(cfg_path(cfg_name)), 1)[0]
    a = arguments_from_file(cfg_path(cfg_name))
    y = measurement(a)
    old = torch.backends.cudnn.allow_tf32
    try:
        for idx in (850, 400):
            x, noise = chain_state(tab, idx, seed=idx)
            for mode, tf32 in (("fp32", False), ("tc", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                operator, cond, sampler = rd.reference_pieces(args, DEV, batch=1)
                xr = x.to(DEV).clone().requires_grad_(True)
                time = torch.tensor([idx], device=DEV)
                out = sampler.p_mean_variance(model=model, x=xr, t=time)
                x0_ref, logvar_ref = out["pred_xstart"].detach().clone(), out["log_variance"].detach().clone()
                freeze = R.utils.is_freeze_phi(args.sample_pattern, idx, sampler.num_timesteps)
                img, loss, vd, grads, aux = cond.conditioning(x_t=out["mean"], measurement=y.to(DEV), noisy_measurement=None, x_prev=xr,
                                                              x_0_hat=out["pred_xstart"], freeze_phi=freeze,
                                                              time_index=float(idx) / sampler.num_timesteps)
                x_next_ref = (img.detach() + torch.exp(0.5 * logvar_ref) * noise.to(DEV)).cpu()
                got = _native_step(mode, cfg_name, idx, x, noise, y)
                want_phi = {n: v.detach().cpu() for n, v in vd.items()}
                assert set(want_phi) == set(got["phi"])
                # against cuDNN-TF32 both sides carry TF32 rounding: same bars as product-vs-oracle
                _check_step(f"{cfg_name} t={idx} [{mode}] vs reference on CUDA (cudnn tf32={tf32})", got, x_next_ref, x0_ref.cpu(),
                            torch.as_tensor(loss), want_phi, mode == "fp32", got["cond"], _amp(tab, idx, x, x0_ref.cpu()))
                assert maxdiff(got["logvar"], logvar_ref.cpu()) < (1e-4 if mode == "fp32" else 2e-2)
    finally:
        torch.backends.cudnn.allow_tf32 = old


# ------------------------------------------------------------------------------------------------- RNG stream on the device


def test_rng_draws_are_bit_identical_to_the_reference_order():
    """gaussian_diffusion.py:149 (dead q_sample draw, randn_like(measurement)) then :266 (randn_like(img)) after
    osmosis_sampling.py:194-197 (manual_seed; randn(x_start_shape, device)) - the FusedStepper's in-place draws consume the
    device generator identically."""
    a, op, cond, sampler = native_objects(CONFIGS[0], 1, respacing=4)
    y = measurement(a).to(DEV)
    torch.manual_seed(a.manual_seed)
    x_T = torch.randn([1, 4, S, S], device=DEV)
    want = []
    for _ in range(2):
        want.append((torch.randn_like(y), torch.randn_like(x_T)))
    torch.manual_seed(a.manual_seed)
    x_T2 = torch.randn([1, 4, S, S], device=DEV)
    stepper = FusedStepper(sampler, full_model("fp32"), cond, x_T2.clone(), y, a.sample_pattern, cuda_graph=False)
    assert torch.equal(x_T, x_T2)
    for k in range(2):
        stepper._draw_into(stepper.dead); stepper._draw_into(stepper.noise)
        assert torch.equal(stepper.dead, want[k][0]) and torch.equal(stepper.noise, want[k][1])


@needs_ref
def test_free_running_loop_from_seed_tracks_the_reference_on_cuda():
    """A 4-step respaced chain from `torch.manual_seed(cfg.manual_seed)` ON THE DEVICE through both public `p_sample_loop`s
    (reference: fp32 convolutions; here: exact mode).  Identical RNG consumption is what makes them agree: the noise term
    is O(1) per pixel, the guidance at most scale * clip = 0.035, so a different draw order would show as O(1) errors."""
    R = rd.import_reference()
    args = R.utils.arguments_from_file(cfg_path(CONFIGS[0]))
    args.diffusion = dict(args.diffusion); args.diffusion["timestep_respacing"] = 4
    model = ref_model()
    a, op, cond, sampler = native_objects(CONFIGS[0], 1, respacing=4)
    y = measurement(a).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        operator, rcond, rsampler = rd.reference_pieces(args, DEV, batch=1)
        torch.manual_seed(args.manual_seed)
        xs = torch.randn([1, 4, S, S], device=DEV).requires_grad_()
        with contextlib.redirect_stderr(open(os.devnull, "w")):
            rimg, rvd, rloss, rx0 = rsampler.p_sample_loop(model=model, x_start=xs, measurement=y, measurement_cond_fn=rcond.conditioning,
                                                           pretrain_model="osmosis", rgb_guidance=False, sample_pattern=args.sample_pattern,
                                                           record=False, save_root=None, image_idx=0, record_every=200,
                                                           original_file_name="t", save_grids_path=None, global_iteration=0)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for seed, same in ((a.manual_seed, True), (a.manual_seed + 1, False)):
        a, op, cond, sampler = native_objects(CONFIGS[0], 1, respacing=4)
        torch.manual_seed(seed)
        xs2 = torch.randn([1, 4, S, S], device=DEV).requires_grad_()
        img, vd, loss, x0 = sampler.p_sample_loop(model=full_model("fp32"), x_start=xs2, measurement=y, measurement_cond_fn=cond.conditioning,
                                                  record=False, save_root=None, pretrain_model="osmosis", rgb_guidance=False,
                                                  sample_pattern=a.sample_pattern, cuda_graph=False)
        torch.cuda.synchronize()
        d = (img.detach().cpu() - rimg.detach().cpu()).abs()
        frac = float((d < 2e-3).float().mean())
        print(f"free-running 4-step chain vs reference on CUDA (seed {seed}): max {float(d.max()):.3e}, within 2e-3: {frac:.4f}")
        if same:
            assert float(d.max()) < 4 * 0.08 and frac > 0.97
            for k in vd:
                assert maxdiff(vd[k].cpu(), rvd[k].detach().cpu()) < 1e-4, k
        else:
            assert frac < 0.5          # another seed: unrelated noise
