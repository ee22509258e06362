"""GPU parity tests of the sampling path (through the package's reference-facing API and the C ABI) against the CPU
oracle and the golden vectors produced by the unmodified reference (tests/golden/).

Tolerances, per level:
  * sampler / guidance kernels: <= 2e-5 normalised (fp32, op-by-op rounding mirrors the reference).
  * UNet, exact mode (fp32 CUDA-core convs): 1e-4 normalised on output and input gradient.
  * UNet, product mode (tcgen05 TF32 convs): 1e-2 normalised - TF32 rounding through ~40 stacked convs; this is the
    arithmetic the reference itself runs on a GPU (cuDNN TF32 convs, SURVEY hazard 6).
  * one guided step (teacher-forced: oracle inputs): x_next within 2 * scale * clip per pixel in TF32 mode (a sign flip of
    a clamped gradient entry), 1e-4 in exact mode; phi to 1e-5.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import osmosis_oracle as orc
from osmosis_diffusion_code_b200 import lib as L_
from osmosis_diffusion_code_b200.guided_diffusion.unet import create_model
from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import create_sampler
from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator, get_noise
from osmosis_diffusion_code_b200.guided_diffusion.condition_methods import get_conditioning_method
from osmosis_diffusion_code_b200.guided_diffusion.posterior_mean_variance import coefficient_table
from tests.golden.cases import CASES, SMALL_UNET, case_inputs
from tests.helpers import golden, small_state_dict, small_cfg, load_yaml_cfg, oracle_specs_from_cfg, rel_err, maxdiff

pytestmark = pytest.mark.gpu
DEV = "cuda"


def lib():
    return L_.load()


def _model(conv_mode):
    m = create_model(**SMALL_UNET, model_path="", conv_mode=conv_mode)
    m.load_state_dict(small_state_dict())
    return m.to(DEV).eval()


_models = {}


def model(conv_mode):
    if conv_mode not in _models:
        _models[conv_mode] = _model(conv_mode)
    return _models[conv_mode]


# ------------------------------------------------------------------------------------------- kernels


def test_timestep_embedding_and_linear_via_unet_time_path():
    # covered end-to-end by test_unet_*; here: the coefficient table equals gather-then-round of the f64 tables
    tab = orc.make_tables(1000, "linear", 250)
    ct = coefficient_table(tab.betas)
    for idx in (0, 1, 100, 249):
        assert ct[idx, 0] == np.float32(tab.sqrt_recip_alphas_cumprod[idx])
        assert ct[idx, 1] == np.float32(tab.sqrt_recipm1_alphas_cumprod[idx])
        assert ct[idx, 2] == np.float32(tab.posterior_mean_coef1[idx])
        assert ct[idx, 3] == np.float32(tab.posterior_mean_coef2[idx])
        assert ct[idx, 4] == np.float32(tab.log_betas[idx])
        assert ct[idx, 5] == np.float32(tab.posterior_log_variance_clipped[idx])


@pytest.mark.parametrize("idx", [0, 3, 999])
def test_posterior_forward_and_vjp(idx):
    tab = orc.make_tables(1000, "linear", 1000)
    g = torch.Generator().manual_seed(idx)
    B, Cc, H, W = 2, 4, 32, 32
    x = torch.randn(B, Cc, H, W, generator=g).requires_grad_(True)
    mo = torch.randn(B, 2 * Cc, H, W, generator=g).requires_grad_(True)
    x0, mean, logvar = orc.posterior(tab, idx, x, mo)
    coef = torch.from_numpy(coefficient_table(tab.betas)).to(DEV)
    t_idx = torch.full((B,), idx, dtype=torch.int32, device=DEV)
    o = [torch.empty(B, Cc, H, W, device=DEV) for _ in range(3)]
    xd, mod = x.detach().to(DEV), mo.detach().to(DEV)  # keep the device copies alive across the async launch
    L_.check(lib().osm_posterior_fwd(L_.ptr(coef), L_.ptr(t_idx), L_.ptr(xd), L_.ptr(mod),
                                     L_.ptr(o[0]), L_.ptr(o[1]), L_.ptr(o[2]), B, Cc, H * W, L_.stream()))
    torch.cuda.synchronize()
    # op-by-op rounding is mirrored: expect bit-exact
    assert maxdiff(o[0].cpu(), x0.detach()) == 0.0
    assert maxdiff(o[1].cpu(), mean.detach()) == 0.0
    assert maxdiff(o[2].cpu(), logvar.detach()) == 0.0
    g0, gm, gl = (torch.randn(B, Cc, H, W, generator=g) for _ in range(3))
    gx_ref, gmo_ref = torch.autograd.grad([x0, mean, logvar], [x, mo], [g0, gm, gl])
    gx = torch.empty(B, Cc, H, W, device=DEV); gmo = torch.empty(B, 2 * Cc, H, W, device=DEV)
    g0d, gmd, gld = g0.to(DEV), gm.to(DEV), gl.to(DEV)
    L_.check(lib().osm_posterior_vjp(L_.ptr(coef), L_.ptr(t_idx), L_.ptr(g0d), L_.ptr(gmd), L_.ptr(gld),
                                     L_.ptr(gx), L_.ptr(gmo), B, Cc, H * W, L_.stream()))
    torch.cuda.synchronize()
    assert rel_err(gx.cpu(), gx_ref) < 2e-6
    assert rel_err(gmo.cpu(), gmo_ref) < 2e-6


def test_sampler_update_and_uncond_update():
    g = torch.Generator().manual_seed(3)
    B, Cc, H, W = 3, 4, 16, 16
    mean, ga, gb, lv, z = (torch.randn(B, Cc, H, W, generator=g) for _ in range(5))
    ga *= 0.01; gb *= 0.01
    scale = torch.tensor([7.0, 7.0, 7.0, 0.9])
    md, gad, gbd, sd, lvd, zd = (v.to(DEV) for v in (mean, ga, gb, scale, lv, z))
    for clip in (0.005, -1.0):
        for tval in (0, 5):
            gsum = ga + gb
            gc = gsum.clamp(-clip, clip) if clip >= 0 else gsum
            want = mean - scale[None, :, None, None] * gc
            if tval != 0:
                want = want + torch.exp(0.5 * lv) * z
            t_idx = torch.full((B,), tval, dtype=torch.int32, device=DEV)
            out = torch.empty(B, Cc, H, W, device=DEV); gout = torch.empty_like(out)
            L_.check(lib().osm_sampler_update(L_.ptr(md), L_.ptr(gad), L_.ptr(gbd), L_.ptr(sd),
                                              clip, L_.ptr(lvd), L_.ptr(zd), L_.ptr(t_idx), L_.ptr(out),
                                              L_.ptr(gout), B, Cc, H * W, L_.stream()))
            torch.cuda.synchronize()
            assert maxdiff(gout.cpu(), gsum) == 0.0
            assert rel_err(out.cpu(), want) < 3e-6   # expf vs the host's exp: a few ulp, and the host's vectorised exp varies by CPU
    x = torch.randn(1, 4, H, W, generator=g); mo = torch.randn(1, 8, H, W, generator=g); zz = torch.randn(1, 4, H, W, generator=g)
    a, ab, bt = 0.98, 0.3, 0.015
    want = orc.ddpm_uncond_update(x, mo[:, :4], zz, np.float64(a), np.float64(ab), np.float64(bt))
    xd = x.to(DEV).clone()
    mod, zzd = mo.to(DEV), zz.to(DEV)
    L_.check(lib().osm_ddpm_uncond_update(L_.ptr(xd), L_.ptr(mod), L_.ptr(zzd), float(1 / np.sqrt(a)),
                                          float((1 - a) / np.sqrt(1 - ab)), float(np.sqrt(bt)), 1, 4, 8, H * W, L_.stream()))
    torch.cuda.synchronize()
    assert rel_err(xd.cpu(), want) < 1e-6


def _native_objects(cname, B):
    c = CASES[cname]
    cfg = load_yaml_cfg(c["yaml"], c["respacing"])
    opcfg = dict(cfg["measurement"]["operator"]); opcfg["batch_size"] = B
    op = get_operator(device=DEV, **opcfg)
    noiser = get_noise(**cfg["measurement"]["noise"])
    cond = get_conditioning_method(cfg["conditioning"]["method"], op, noiser, **cfg["conditioning"]["params"],
                                   **cfg["sample_pattern"], **cfg["aux_loss"])
    sampler = create_sampler(**cfg["diffusion"])
    return cfg, op, cond, sampler


@pytest.mark.parametrize("size", [32, 96], ids=["1cta", "cluster8"])
@pytest.mark.parametrize("cname", list(CASES))
def test_operator_forward_and_guidance_loop(cname, size):
    """osm_operator_forward and osm_guidance_phi_loop vs the oracle's autograd version, B=3 (per-image semantics).
    size 32 runs one CTA per image, size 96 the 8-CTA cluster path with the DSMEM reduction."""
    B = 3
    cfg, op, cond, sampler = _native_objects(cname, B)
    tab, ospec, gspec, phis, names = oracle_specs_from_cfg(cfg, B)
    g = torch.Generator().manual_seed(17)
    y1, xgt = case_inputs("meas:" + cname)
    if size != y1.shape[-1]:
        y1 = torch.nn.functional.interpolate(y1, size=size, mode="bilinear", align_corners=False)
        xgt = torch.nn.functional.interpolate(xgt, size=size, mode="bilinear", align_corners=False)
    H = y1.shape[-1]
    x0 = (xgt + 0.3 * torch.randn(B, 4, H, H, generator=g)).contiguous()
    y = (y1 + 0.05 * torch.randn(B, 3, H, H, generator=g)).contiguous()
    assert rel_err(op.forward(x0.to(DEV)).cpu(), orc.operator_forward(ospec, x0, phis)) < 2e-6
    for freeze in (True, False):
        cfgB, opB, condB, _ = _native_objects(cname, B)
        # oracle: n evaluations with SGD after each, x-gradient at the last
        ph = [p.clone() for p in phis]
        n = 1 if freeze else gspec.n_iter
        for it in range(n):
            phr = [p.clone().requires_grad_(not freeze) for p in ph]
            x0r = x0.clone().requires_grad_(it == n - 1)
            total, norm, terms = orc.guidance_losses(ospec, x0r, y, phr, gspec.loss_weight, gspec.weight_fn, gspec.aux)
            wrt = ([x0r] if it == n - 1 else []) + (phr if not freeze else [])
            grads = torch.autograd.grad(total.sum(), wrt)
            if it == n - 1:
                gx0_ref, grads = grads[0], grads[1:]
            if not freeze:
                ph = [(p.detach() - eta * gp) for p, eta, gp in zip(phr, ospec.eta, grads)]
        fz = torch.tensor([1 if freeze else 0], dtype=torch.int32, device=DEV)
        gx0 = torch.empty(B, 4, H, H, device=DEV); losses = torch.zeros(B, 4, device=DEV)
        x0d, yd = x0.to(DEV), y.to(DEV)
        condB.guidance_gradient(x0d, yd, fz, gx0, losses)
        torch.cuda.synchronize()
        assert rel_err(losses[:, 0].cpu(), norm.detach()) < 2e-6
        assert rel_err(gx0.cpu(), gx0_ref) < 2e-5
        for nme, p in zip(names, ph):
            assert maxdiff(getattr(opB, nme).cpu(), p) < 2e-6, nme


# ------------------------------------------------------------------------------------------- UNet


@pytest.mark.parametrize("conv_mode,tol", [("fp32", 1e-4), ("tc", 1e-2)])
def test_unet_forward_and_input_vjp_vs_reference_golden(conv_mode, tol):
    gold = golden()
    x, t, cot = case_inputs("unet")
    m = model(conv_mode)
    xd = x.to(DEV).requires_grad_(True)
    out = m(xd, t.to(DEV))
    (gx,) = torch.autograd.grad(out, xd, cot.to(DEV))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and torch.isfinite(gx).all()
    assert rel_err(out.detach().cpu(), gold["unet_out"]) < tol
    assert rel_err(gx.cpu(), gold["unet_gx"]) < tol


def test_unet_determinism_and_batch_shard_invariance():
    """Same batch size twice -> bit-identical (fixed-order reductions, no float atomics).  B images at once vs one by one:
    identical up to summation order (and the TF32 operand roundings it can flip) - the conv tile / split-K policy depends on the per-GPU batch (with equal
    per-GPU batches, as in the weak-scaling runs, every image sees the same policy on every rank)."""
    m = model("tc")
    x, t, cot = case_inputs("unet")
    xd, td, cd = x.to(DEV), t.to(DEV), cot.to(DEV)
    full = m._forward_raw(xd, td.float()).clone()
    gfull = m._vjp_raw(cd).clone()
    full2 = m._forward_raw(xd, td.float()).clone()
    gfull2 = m._vjp_raw(cd).clone()
    assert torch.equal(full, full2) and torch.equal(gfull, gfull2)
    for b in range(x.shape[0]):
        one = m._forward_raw(xd[b:b + 1], td[b:b + 1].float())
        gone = m._vjp_raw(cd[b:b + 1].contiguous())
        assert rel_err(one[0].cpu(), full[b].cpu()) < 2e-3 and rel_err(gone[0].cpu(), gfull[b].cpu()) < 2e-3  # TF32-level


# ------------------------------------------------------------------------------------------- steps and loop


@pytest.mark.parametrize("conv_mode", ["fp32", "tc"])
@pytest.mark.parametrize("cname", list(CASES))
def test_single_guided_step_vs_reference_golden(cname, conv_mode):
    gold, c = golden(), CASES[cname]
    m = model(conv_mode)
    for idx in c["step_idx"]:
        cfg, op, cond, sampler = _native_objects(cname, 1)
        y, _ = case_inputs("meas:" + cname)
        x = case_inputs(f"x:{cname}:{idx}").to(DEV)
        noise = case_inputs(f"noise:{cname}:{idx}").to(DEV)
        st = sampler.fused_state(m, cond, x, y.to(DEV))
        from osmosis_diffusion_code_b200.osmosis_utils.utils import is_freeze_phi
        freeze = is_freeze_phi(cfg["sample_pattern"], idx, sampler.num_timesteps)
        st["t_idx"].fill_(idx); st["t_model"].fill_(sampler._model_timestep(idx)); st["freeze"].fill_(int(freeze))
        img = x.clone()
        sampler.fused_step(m, cond, st, img, noise)
        torch.cuda.synchronize()
        pre = f"{cname}/step{idx}/"
        assert bool(gold[pre + "freeze"][0]) == freeze
        exact = conv_mode == "fp32"
        # The first step of the chain (t = 999: sqrt(1/abar) ~ sqrt(1/abar - 1) ~ 157) is ill-conditioned: x0_hat =
        # 157 (x - eps) amplifies the UNet's rounding ~157x before the operator's exp(-phi d); the loss there is ~1e7..1e14
        # and dominated by a few pixels (SURVEY hazard 2).  Its value is compared loosely; what the step does with it
        # (the clamped update, phi) is still checked below.
        first = idx == sampler.num_timesteps - 1
        assert rel_err(st["losses"][:, 0].cpu(), gold[pre + "loss"]) < ((5e-3 if exact else 0.3) if first else (1e-4 if exact else 1e-2))
        assert rel_err(st["grad"].cpu(), gold[pre + "grad"]) < ((2e-2 if exact else 0.5) if first else (1e-3 if exact else 5e-2))
        clipv = cond.gradient_clip_value
        bound = 2 * float(cond.scale.max()) * clipv * 1.01 + 1e-2 * float(np.abs(gold[pre + "x_next"]).max()) * (0.0 if exact else 1.0)
        d = (img.cpu() - torch.from_numpy(gold[pre + "x_next"])).abs()
        assert float(d.max()) <= (bound if not exact else max(bound, 1e-4))
        # almost every pixel agrees tightly; only clamp-boundary sign flips may differ
        frac_tight = float((d < (2e-4 if exact else 2e-2) * max(1.0, float(np.abs(gold[pre + "x_next"]).max()))).float().mean())
        assert frac_tight > (0.999 if exact else 0.98)
        for n in op.groups:
            assert maxdiff(getattr(op, n).cpu(), gold[pre + n]) < (2e-6 if exact else 2e-5), n


@pytest.mark.parametrize("cname", list(CASES))
def test_loop_paths_agree_and_track_reference(cname):
    """p_sample_loop through the reference-facing API: fused path == autograd-compatible path (same kernels), and the
    exact-mode chain tracks the reference's own 6-step p_sample_loop output (free-running, statistical bound)."""
    gold, c = golden(), CASES[cname]
    y, _ = case_inputs("meas:" + cname)
    res = {}
    for fused in (True, False):
        cfg, op, cond, sampler = _native_objects(cname, 1)
        m = model("fp32")
        torch.manual_seed(cfg["manual_seed"])
        x_start = torch.randn(1, 4, *y.shape[2:], device=DEV).requires_grad_()
        img, vd, loss, x0 = sampler.p_sample_loop(model=m, x_start=x_start, measurement=y.to(DEV),
                                                  measurement_cond_fn=cond.conditioning, record=False, save_root=None,
                                                  pretrain_model="osmosis", rgb_guidance=False,
                                                  sample_pattern=cfg["sample_pattern"], fused=fused)
        torch.cuda.synchronize()
        res[fused] = (img.detach().cpu(), {k: v.cpu() for k, v in vd.items()}, loss, x0)
    assert maxdiff(res[True][0], res[False][0]) < 1e-5
    assert maxdiff(res[True][3], res[False][3]) < 1e-5
    for k in res[True][1]:
        assert maxdiff(res[True][1][k], res[False][1][k]) < 1e-6
    assert x0.device.type == "cpu" and tuple(res[True][1][list(res[True][1])[0]].shape[2:]) == (1, 1)


def test_loop_teacher_forced_chain_vs_oracle():
    """6-step chain with the oracle's noise injected; compared step-free at the end in exact mode.  The GPU RNG differs
    from the CPU RNG the golden loop used, so the chain is re-run on the oracle with the SAME noise tensors."""
    cname = "osmosis"
    c = CASES[cname]
    cfg, op, cond, sampler = _native_objects(cname, 1)
    tab, ospec, gspec, phis, names = oracle_specs_from_cfg(cfg, 1)
    y, _ = case_inputs("meas:" + cname)
    g = torch.Generator().manual_seed(99)
    x_T = torch.randn(1, 4, *y.shape[2:], generator=g)
    noises = {idx: torch.randn(1, 4, *y.shape[2:], generator=g) for idx in range(tab.num_timesteps)}
    xo, pho, losso, x0o = orc.sample_loop(small_state_dict(), small_cfg(), tab, ospec, gspec, x_T, y, phis, lambda i: noises[i])
    m = model("fp32")
    img = x_T.to(DEV).clone()
    st = sampler.fused_state(m, cond, img, y.to(DEV))
    from osmosis_diffusion_code_b200.osmosis_utils.utils import is_freeze_phi
    for idx in range(sampler.num_timesteps)[::-1]:
        st["t_idx"].fill_(idx); st["t_model"].fill_(sampler._model_timestep(idx))
        st["freeze"].fill_(int(is_freeze_phi(cfg["sample_pattern"], idx, sampler.num_timesteps)))
        sampler.fused_step(m, cond, st, img, noises[idx].to(DEV))
    torch.cuda.synchronize()
    d = (img.cpu() - xo).abs()
    assert float((d < 1e-3).float().mean()) > 0.995     # clamp sign flips are rare over 6 steps
    assert float(d.max()) < 0.2
    for n, p in zip(names, pho):
        assert maxdiff(getattr(op, n).cpu(), p) < 1e-4


def test_loop_guidance_window_and_alternate_length_vs_oracle():
    """sample_pattern with a guidance window (unguided steps at both ends, gaussian_diffusion.py:218-222, :262-264) and an
    alternate length local_M = 2 inside [s_end, s_start] (utils.py:593-630): the fused loop (FusedStepper, with and without
    CUDA-graph replay) against the oracle's loop with the same injected noise, exact mode."""
    cname = "osmosis"
    cfg, op, cond, sampler = _native_objects(cname, 1)
    sp = cfg["sample_pattern"]
    sp.update(start_guidance=0.8, stop_guidance=0.2, update_start=0.7, update_end=0.2, s_start=0.5, s_end=0.3, local_M=2)
    T = sampler.num_timesteps
    tab, ospec, gspec, phis, names = oracle_specs_from_cfg(cfg, 1)
    assert [orc.alternate_length(gspec, i, T) for i in range(T)].count(2) >= 1
    assert not orc.guidance_on(gspec, T - 1, T) and not orc.guidance_on(gspec, 0, T)
    y, _ = case_inputs("meas:" + cname)
    g = torch.Generator().manual_seed(5)
    x_T = torch.randn(1, 4, *y.shape[2:], generator=g)
    draws = [torch.randn(1, 4, *y.shape[2:], generator=g) for _ in range(3 * T)]
    it = iter(draws)
    xo, pho, losso, x0o = orc.sample_loop(small_state_dict(), small_cfg(), tab, ospec, gspec, x_T, y, phis, lambda i: next(it))
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import FusedStepper
    from osmosis_diffusion_code_b200.osmosis_utils.utils import set_alternate_length
    m = model("fp32")
    for use_graph in (False, True):
        cfg, op, cond, sampler = _native_objects(cname, 1)
        cfg["sample_pattern"].update(sp)
        img = x_T.to(DEV).clone()
        stepper = FusedStepper(sampler, m, cond, img, y.to(DEV), cfg["sample_pattern"], cuda_graph=use_graph)
        k = 0
        n_steps = 0
        for idx in range(T)[::-1]:
            for _ in range(set_alternate_length(cfg["sample_pattern"], idx, T)):
                stepper._draw_into = lambda buf, _k=k: buf.copy_(draws[_k].to(DEV)) if buf.shape[1] == 4 else buf.zero_()
                stepper.step(idx)
                k += 1
                n_steps += 1
        torch.cuda.synchronize()
        assert n_steps == T + [orc.alternate_length(gspec, i, T) for i in range(T)].count(2)
        d = (img.cpu() - xo).abs()
        assert float((d < 1e-3).float().mean()) > 0.995, use_graph
        assert float(d.max()) < 0.2
        for n, p in zip(names, pho):
            assert maxdiff(getattr(op, n).cpu(), p) < 1e-4


def test_loop_reuse_and_host_measurement_progress():
    """Consecutive p_sample_loop calls on one sampler reuse the captured step graph (the reference's per-image loop,
    osmosis_sampling.py:117-199): the second call must give the bits of a fresh sampler.  The measurement comes from host
    memory and the progress callback sees every step once, in order, with the loss the loop finally returns."""
    cname = "osmosis"
    y, _ = case_inputs("meas:" + cname)
    m = model("fp32")

    def run(sampler, cfg, cond, seen):
        torch.manual_seed(cfg["manual_seed"])
        x_start = torch.randn(1, 4, *y.shape[2:], device=DEV)
        out = sampler.p_sample_loop(model=m, x_start=x_start, measurement=y.clone(), measurement_cond_fn=cond.conditioning,
                                    record=False, save_root=None, pretrain_model="osmosis", rgb_guidance=False,
                                    sample_pattern=cfg["sample_pattern"], progress=lambda idx, loss: seen.append((idx, loss.copy())))
        torch.cuda.synchronize()
        return out

    cfg, op, cond, sampler = _native_objects(cname, 1)
    phi0 = {n: getattr(op, n).clone() for n in op.groups}
    seen1, seen2 = [], []
    img1, vd1, loss1, x01 = run(sampler, cfg, cond, seen1)
    for n in op.groups:                       # same operator object, phi restored in place: the graph's pointers stay valid
        getattr(op, n).data.copy_(phi0[n])
    img2, vd2, loss2, x02 = run(sampler, cfg, cond, seen2)
    assert len(sampler._steppers) == 1
    assert torch.equal(img1, img2) and torch.equal(x01, x02) and np.array_equal(loss1, loss2)
    T = sampler.num_timesteps
    assert [i for i, _ in seen1] == list(range(T))[::-1] == [i for i, _ in seen2]
    assert np.array_equal(seen1[-1][1], loss1)
    for (_, a), (_, b) in zip(seen1, seen2):
        assert np.array_equal(a, b)


def test_simulation_batch_psnr_vs_oracle():
    """BASELINE config 4 in miniature (osmosis_simulation_sample_config, a batch of distinct synthetic scenes, product mode =
    tcgen05 TF32 convs + fused attention): PSNR of the restored RGB (pred_xstart[:, :3], data range 2.0) against the oracle's
    fp32 chain with the same injected noise, per image, after the 6-step chain.  Also the per-image semantics of a batch: the
    batched run equals the images run one by one to TF32 rounding."""
    from osmosis_diffusion_code_b200.synthetic import synth_measurement
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import FusedStepper
    cname, B = "simulation", 3
    cfg, op, cond, sampler = _native_objects(cname, B)
    tab, ospec, gspec, phis, names = oracle_specs_from_cfg(cfg, B)
    ys = torch.cat([synth_measurement(10 + i, 32, phi_a=(1.1, 0.95, 0.95), phi_b=(1.1, 0.95, 0.95), phi_inf=(0.2, 0.4, 0.7),
                                      depth_type="original")[0] for i in range(B)], 0)
    g = torch.Generator().manual_seed(21)
    x_T = torch.randn(1, 4, 32, 32, generator=g).repeat(B, 1, 1, 1)           # the reference reseeds per image: shared x_T / noise
    noises = {idx: torch.randn(1, 4, 32, 32, generator=g).repeat(B, 1, 1, 1) for idx in range(tab.num_timesteps)}
    xo, pho, losso, x0o = orc.sample_loop(small_state_dict(), small_cfg(), tab, ospec, gspec, x_T, ys, phis, lambda i: noises[i])

    def run(model_, cond_, sampler_, y_, x_, nz):
        img = x_.to(DEV).clone()
        stepper = FusedStepper(sampler_, model_, cond_, img, y_.to(DEV), cfg["sample_pattern"], cuda_graph=True)
        for idx in range(sampler_.num_timesteps)[::-1]:
            stepper._draw_into = lambda buf, _i=idx: buf.copy_(nz[_i].to(DEV)) if buf.shape[1] == 4 else buf.zero_()
            stepper.step(idx)
        torch.cuda.synchronize()
        return stepper.st["x0"].cpu()

    x0 = run(model("tc"), cond, sampler, ys, x_T, noises)
    mse = ((x0[:, :3] - x0o[:, :3]) ** 2).mean(dim=(1, 2, 3))
    psnr = 10 * torch.log10(4.0 / mse)
    assert float(psnr.min()) > 40.0, psnr          # measured 50-60 dB: TF32 rounding + rare clamp sign flips over 6 steps
    for b in range(B):
        _, op1, cond1, sampler1 = _native_objects(cname, 1)
        x0_1 = run(model("tc"), cond1, sampler1, ys[b:b + 1], x_T[b:b + 1], {k: v[b:b + 1] for k, v in noises.items()})
        m1 = ((x0_1[:, :3] - x0[b:b + 1, :3]) ** 2).mean()
        assert float(10 * torch.log10(4.0 / m1.clamp_min(1e-20))) > 40.0


def test_uncond_inverse_vs_reference_golden(tmp_path, monkeypatch):
    """BASELINE config 1 (`RGBD_prior_sampling.py`): `osmosis_utils.diffusion.GaussianDiffusion.inverse` with the native UNet and
    the osm_ddpm_uncond_update kernel against the unmodified reference's run (CPU RNG draws injected), including the recorded
    x_start_rgb / depth outputs and the process grid file."""
    import os
    from osmosis_diffusion_code_b200.osmosis_utils.diffusion import GaussianDiffusion
    from osmosis_diffusion_code_b200.osmosis_utils import utils as utilso
    from tests.golden.cases import UNCOND_CASE, SMALL_HW
    gold = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "uncond_golden.npz")))
    c = UNCOND_CASE
    torch.manual_seed(c["seed"])
    x_T = torch.randn(1, 4, SMALL_HW, SMALL_HW)
    zs = [torch.randn(1, 4, SMALL_HW, SMALL_HW) for _ in range(c["steps"] - 1)]
    it = iter(zs)
    monkeypatch.setattr(torch, "randn_like", lambda t, **kw: next(it).to(t.device))
    monkeypatch.setattr(utilso, "depth_tensor_to_color_image", lambda d, **kw: d.repeat(3, 1, 1) if d.dim() == 3 else d)  # identity map
    m = model("fp32")
    x, (rgb, depth) = GaussianDiffusion(T=c["T"], schedule="linear").inverse(
        net=m, shape=(4, SMALL_HW, SMALL_HW), image_channels=4, steps=c["steps"], start_t=c["start_t"], device=DEV, x=x_T.to(DEV),
        record_process=True, record_every=200, save_path=str(tmp_path), image_idx=0)
    torch.cuda.synchronize()
    assert maxdiff(x.cpu(), gold["x"]) < 1e-4 * max(1.0, float(np.abs(gold["x"]).max()))
    assert maxdiff(rgb.cpu(), gold["x_start_rgb"]) < 1e-4
    assert maxdiff(depth[0:1].cpu(), gold["x_depth_pmm"]) < 5e-4
    assert os.path.exists(os.path.join(str(tmp_path), "image_0_process.png"))


def test_stepper_graph_is_dropped_when_the_model_is_rebound():
    """A cached FusedStepper's CUDA graph holds pointers into the engine's bound workspace; running another batch shape
    through the same model re-plans (and frees) it.  The stepper must notice (UNetModel._bind_generation) and re-capture
    instead of replaying the stale plan: results equal those of an undisturbed stepper bit for bit."""
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import FusedStepper
    cname = "osmosis"
    y, _ = case_inputs("meas:" + cname)
    m = model("fp32")
    g = torch.Generator().manual_seed(31)
    x_T = torch.randn(1, 4, *y.shape[2:], generator=g)
    T = 6
    noises = [torch.randn(1, 4, *y.shape[2:], generator=g) for _ in range(T)]

    def chain(disturb):
        cfg, op, cond, sampler = _native_objects(cname, 1)
        img = x_T.to(DEV).clone()
        stepper = FusedStepper(sampler, m, cond, img, y.to(DEV), cfg["sample_pattern"], cuda_graph=True)
        for k, idx in enumerate(range(sampler.num_timesteps)[::-1][:T]):
            if disturb and k == 3:     # graph captured at k == 1; now another (B, H, W) goes through the same model
                gen = m._bind_generation
                m._forward_raw(torch.randn(2, 4, 16, 16, device=DEV), torch.full((2,), 10.0, device=DEV))
                assert m._bind_generation == gen + 1
            stepper._draw_into = lambda buf, _k=k: buf.copy_(noises[_k].to(DEV)) if buf.shape[1] == 4 else buf.zero_()
            stepper.step(idx)
        torch.cuda.synchronize()
        return img.clone(), op.phi.clone(), stepper

    a_img, a_phi, _ = chain(False)
    b_img, b_phi, st = chain(True)
    assert st.graph is not None                       # re-captured after the disturbance
    assert torch.equal(a_img, b_img) and torch.equal(a_phi, b_phi)


def test_batch_mismatch_and_unknown_operator_raise():
    """phi / optimizer state hold one row per image (`batch_size` of get_operator): running another batch size must fail
    loudly instead of indexing out of bounds - on the host (ValueError) and in the C ABI (phi_batch != B).  An operator
    the guidance kernel does not know raises NotImplementedError instead of a KeyError."""
    cname = "osmosis"
    cfg, op, cond, sampler = _native_objects(cname, 1)
    y, _ = case_inputs("meas:" + cname)
    x2 = torch.randn(2, 4, *y.shape[2:], device=DEV)
    y2 = y.to(DEV).repeat(2, 1, 1, 1)
    with pytest.raises(ValueError):
        sampler.fused_state(model("fp32"), cond, x2, y2)
    with pytest.raises(ValueError):
        cond.conditioning(x_prev=x2, x_t=x2.clone(), x_0_hat=x2.clone(), measurement=y2, freeze_phi=True)
    from osmosis_diffusion_code_b200.osmosis_utils.utils import postprocess_samples
    with pytest.raises(ValueError):
        postprocess_samples(op, x2, y2)
    p = cond.kernel_params()
    assert p.phi_batch == 1
    buf = dict(f=torch.zeros(1, dtype=torch.int32, device=DEV), g=torch.empty_like(x2), l=torch.zeros(2, 4, device=DEV))
    rc = lib().osm_guidance_phi_loop(C.byref(p), L_.ptr(x2), L_.ptr(y2), L_.ptr(op.phi), L_.ptr(buf["f"]), L_.ptr(buf["g"]), L_.ptr(buf["l"]),
                                     2, x2.shape[2] * x2.shape[3], L_.stream())
    assert rc != 0 and b"phi_batch" in lib().osm_last_error_string()

    class Foreign:
        kind_name = "my_new_operator"
        phi = op.phi
    cond2 = get_conditioning_method(cfg["conditioning"]["method"], Foreign(), get_noise(**cfg["measurement"]["noise"]),
                                    **cfg["conditioning"]["params"], **cfg["sample_pattern"], **cfg["aux_loss"])
    with pytest.raises(NotImplementedError):
        cond2.kernel_params()


def test_unet_with_groupnorm_fused_into_the_conv_operand_vs_reference_golden(monkeypatch):
    """OSM_GN_XFORM=2 forces the halo conv kernel with the in-shared-memory GroupNorm + SiLU transform onto every 3x3 conv
    whose shape allows it (here: all non-resampling ResBlock convs of the small golden UNet; at full size the planner
    picks it for the 256x256 and 128x128 levels, which tests/test_fullsize_parity_gpu.py covers).  Output and input-VJP
    against the unmodified reference's golden, product-mode tolerance; and the launch count drops (no apply passes)."""
    gold = golden()
    x, t, cot = case_inputs("unet")
    base = model("tc")
    base._forward_raw(x.to(DEV), t.to(DEV).float())
    n_base = base.launch_counts()[0]
    monkeypatch.setenv("OSM_GN_XFORM", "2")
    m = _model("tc")
    xd = x.to(DEV).requires_grad_(True)
    out = m(xd, t.to(DEV))
    (gx,) = torch.autograd.grad(out, xd, cot.to(DEV))
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu(), gold["unet_out"]) < 1e-2
    assert rel_err(gx.cpu(), gold["unet_gx"]) < 1e-2
    kinds = [o["kind"] for o in m.profile_ops(0)]
    kinds_base = [o["kind"] for o in base.profile_ops(0)]
    assert kinds.count("gn_apply") < kinds_base.count("gn_apply")
    print("forward launches with / without the fused operand transform:", m.launch_counts()[0], n_base,
          "gn_apply ops", kinds.count("gn_apply"), "vs", kinds_base.count("gn_apply"))


def test_unet_with_fp16_operand_convs_vs_reference_golden(monkeypatch):
    """OSM_CONV_F16=2 puts every 3x3 conv whose shape allows it (forward and input-gradient, with the GroupNorm + SiLU operand
    transform, the fused forward / backward statistics epilogues and the power-of-two prescale of the cotangent) on the
    fp16-operand tcgen05 kernel (conv_tc_halo16_2sm_kernel).  fp16 and TF32 carry the same 11-bit significand, so output and
    input-VJP must meet the SAME product-mode tolerance against the unmodified reference's golden as the TF32 kernels, and
    the error must not be larger than theirs by more than rounding noise; a cotangent scaled down to 1e-9 of its size (far
    below fp16's range) must give exactly 1e-9 of the gradient's accuracy class (the prescale makes the VJP scale-free)."""
    gold = golden()
    x, t, cot = case_inputs("unet")
    monkeypatch.setenv("OSM_CONV_F16", "0")
    m32 = _model("tc")
    monkeypatch.setenv("OSM_CONV_F16", "2")
    m16 = _model("tc")
    kinds = {}
    errs = {}
    for name, m in (("tf32", m32), ("fp16", m16)):
        xd = x.to(DEV).requires_grad_(True)
        out = m(xd, t.to(DEV))
        (gx,) = torch.autograd.grad(out, xd, cot.to(DEV))
        torch.cuda.synchronize()
        errs[name] = (rel_err(out.detach().cpu(), gold["unet_out"]), rel_err(gx.cpu(), gold["unet_gx"]))
        kinds[name] = m.launch_counts()
    print("UNet error vs the reference golden (output, input-VJP): tf32 convs", errs["tf32"], "fp16-operand convs", errs["fp16"],
          "launches", kinds)
    assert errs["fp16"][0] < 1e-2 and errs["fp16"][1] < 1e-2
    assert errs["fp16"][0] < 2 * errs["tf32"][0] + 1e-4 and errs["fp16"][1] < 2 * errs["tf32"][1] + 1e-4
    # scale invariance of the input-VJP: tiny and huge cotangents
    xd = x.to(DEV).requires_grad_(True)
    out = m16(xd, t.to(DEV))
    (g1,) = torch.autograd.grad(out, xd, cot.to(DEV), retain_graph=True)
    for s in (2.0 ** -30, 2.0 ** 20):
        xs = x.to(DEV).requires_grad_(True)
        (g2,) = torch.autograd.grad(m16(xs, t.to(DEV)), xs, (cot * s).to(DEV))
        assert torch.equal(g2 / s, g1), f"the input-VJP is not exactly linear under a power-of-two scale of {s}"
