"""GPU, BASELINE sizes: the shipped config's 553 M-parameter UNet at 256x256 (too large for the CPU oracle to finish in
seconds), checked through size-independent properties:

  * product mode (tcgen05 TF32 convs, CTA-pair kernel, fused tcgen05 attention, fused GroupNorm statistics) against exact
    mode (fp32 CUDA-core convs, fp32-accurate attention) on the same weights: UNet output and input-VJP agree to TF32
    rounding (1e-2 / 3e-2 normalised - the arithmetic the reference itself runs on a GPU, SURVEY hazard 6);
  * the hand-written backward is the adjoint of the forward: <J v, g> from central finite differences of the exact-mode
    forward equals <v, J^T g> from osm_unet_vjp_input; and it is linear in g;
  * bit-reproducibility run to run; a batch of 2 equals two single images to TF32 rounding (batch-shard semantics);
  * one full guided step of each shipped guided config (all three operators): finite, the update bounded by
    scale * clip per pixel, phi moves only when it is not frozen.
"""
import contextlib
import os
import sys

import pytest
import torch

from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file, is_freeze_phi
from osmosis_diffusion_code_b200.guided_diffusion.unet import create_model
from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import create_sampler
from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator, get_noise
from osmosis_diffusion_code_b200.guided_diffusion.condition_methods import get_conditioning_method
from osmosis_diffusion_code_b200.synthetic import synth_state_dict, synth_measurement
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV, S = "cuda", 256
_cache = {}


def full_model(conv_mode):
    if conv_mode not in _cache:
        a = arguments_from_file(os.path.join(ROOT, "configs", "osmosis_sample_config.yaml"))
        um = dict(a.unet_model); um["model_path"] = ""
        with contextlib.redirect_stdout(sys.stderr):
            m = create_model(**um, conv_mode=conv_mode)
        sd = synth_state_dict(m.param_specs(), um["num_channels"], seed=7, delta=0.05)
        m.load_state_dict(sd); del sd
        _cache[conv_mode] = m.to(DEV).eval()
    return _cache[conv_mode]


def inputs(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 4, S, S, generator=g).to(DEV)
    t = torch.full((B,), 500.0, device=DEV)
    cot = (torch.randn(B, 8, S, S, generator=g) * 1e-3).to(DEV)
    return x, t, cot


def test_product_mode_matches_exact_mode_at_full_size():
    x, t, cot = inputs(1)
    outs = {}
    for mode in ("fp32", "tc"):
        m = full_model(mode)
        assert m.num_params() == 552_821_000 or m.num_params() > 5.5e8
        o = m._forward_raw(x, t).clone()
        g = m._vjp_raw(cot).clone()
        o2 = m._forward_raw(x, t).clone()
        g2 = m._vjp_raw(cot).clone()
        torch.cuda.synchronize()
        assert torch.isfinite(o).all() and torch.isfinite(g).all()
        assert torch.equal(o, o2) and torch.equal(g, g2)                      # bit-reproducible (no float atomics)
        outs[mode] = (o.cpu(), g.cpu())
    assert rel_err(outs["tc"][0], outs["fp32"][0]) < 1e-2
    assert rel_err(outs["tc"][1], outs["fp32"][1]) < 3e-2


def test_vjp_is_the_adjoint_of_the_forward_and_linear():
    m = full_model("fp32")
    x, t, cot = inputs(1, seed=3)
    g = torch.Generator().manual_seed(9)
    v = torch.randn(1, 4, S, S, generator=g).to(DEV)
    cot2 = (torch.randn(1, 8, S, S, generator=g) * 1e-3).to(DEV)
    eps = 1e-2
    fp = m._forward_raw(x + eps * v, t).double().clone()
    fm = m._forward_raw(x - eps * v, t).double().clone()
    jv = (fp - fm) / (2 * eps)
    # cotangent along J v itself: <J v, g> is then a sum of squares (no cancellation), so the finite-difference noise of the
    # fp32 forward (~1e-4 relative) is all that separates the two sides
    cot = (jv / jv.abs().max() * 1e-3).float().contiguous()
    m._forward_raw(x, t)
    jt_g = m._vjp_raw(cot).double().clone()
    lhs, rhs = float((jv * cot.double()).sum()), float((v.double() * jt_g).sum())
    assert lhs > 0 and abs(lhs - rhs) <= 3e-3 * abs(lhs), (lhs, rhs)
    m._forward_raw(x, t)
    jt_g2 = m._vjp_raw(cot2).double().clone()
    m._forward_raw(x, t)
    both = m._vjp_raw((0.5 * cot - 2.0 * cot2).contiguous()).double()
    assert rel_err(both.cpu(), (0.5 * jt_g - 2.0 * jt_g2).cpu()) < 1e-4


def test_batch_of_two_equals_two_single_images():
    m = full_model("tc")
    x, t, cot = inputs(2, seed=5)
    o = m._forward_raw(x, t).clone()
    g = m._vjp_raw(cot).clone()
    for b in range(2):
        ob = m._forward_raw(x[b:b + 1].contiguous(), t[b:b + 1]).clone()
        gb = m._vjp_raw(cot[b:b + 1].contiguous()).clone()
        assert rel_err(ob[0].cpu(), o[b].cpu()) < 3e-3 and rel_err(gb[0].cpu(), g[b].cpu()) < 1e-2     # tile policy differs: TF32-level


@pytest.mark.parametrize("cfg_name", ["osmosis_sample_config.yaml", "osmosis_simulation_sample_config.yaml", "osmosis_haze_sample_config.yaml"])
def test_full_guided_step_of_each_config(cfg_name):
    a = arguments_from_file(os.path.join(ROOT, "configs", cfg_name))
    m = full_model("tc")
    B = 2
    opc = dict(a.measurement["operator"]); opc["batch_size"] = B
    op = get_operator(device=DEV, **opc)
    cond = get_conditioning_method(a.conditioning["method"], op, get_noise(**a.measurement["noise"]), **a.conditioning["params"],
                                   **a.sample_pattern, **a.aux_loss)
    d = dict(a.diffusion); d["timestep_respacing"] = 1000
    sampler = create_sampler(**d)
    ph = lambda k, dflt: [float(v) for v in str(opc.get(k, dflt)).split(",")]
    pa, pb = (ph("phi_a", "1"), ph("phi_b", "1")) if "phi_a" in opc else (ph("phi_ab", "1"), ph("phi_ab", "1"))
    y = torch.cat([synth_measurement(i, S, pa, pb, ph("phi_inf", "0.2,0.4,0.7"), depth_type=opc.get("depth_type"))[0] for i in range(B)]).to(DEV)
    g = torch.Generator().manual_seed(1)
    for idx in (900, 400):                                     # a frozen-phi step and an optimised-phi step (20 inner iterations)
        x = torch.randn(B, 4, S, S, generator=g).to(DEV) * (1.0 if idx > 500 else 0.6)
        noise = torch.randn(B, 4, S, S, generator=g).to(DEV)
        img = x.clone()
        st = sampler.fused_state(m, cond, img, y)
        freeze = is_freeze_phi(a.sample_pattern, idx, sampler.num_timesteps)
        st["t_idx"].fill_(idx); st["t_model"].fill_(sampler._model_timestep(idx)); st["freeze"].fill_(int(freeze))
        phi_before = op.phi.clone()
        sampler.fused_step(m, cond, st, img, noise)
        torch.cuda.synchronize()
        assert torch.isfinite(img).all() and torch.isfinite(st["losses"]).all() and torch.isfinite(op.phi).all()
        unguided = st["mean"] + torch.exp(0.5 * st["logvar"]) * noise
        bound = cond._scale4(4).to(DEV)[None, :, None, None] * cond.gradient_clip_value
        assert bool(((img - unguided).abs() <= bound * 1.001 + 1e-5).all())
        assert float((img - unguided).abs().max()) > 0                      # guidance did act
        moved = float((op.phi - phi_before).abs().max())
        assert (moved == 0.0) if freeze else (moved > 0.0)
        assert float((st["losses"][0] - st["losses"][1]).abs().max()) > 0    # distinct images -> distinct per-image losses


def test_full_size_chain_psnr_product_vs_exact_mode():
    """BASELINE config 4's check at full size (osmosis_simulation_sample_config, 256x256, 553 M parameters, a batch of two
    synthetic scenes): a 12-step respaced guided chain in product mode (tcgen05 TF32 convs / fused attention / fused GroupNorm
    statistics) against the same chain in exact mode (fp32 CUDA-core convs, fp32-accurate attention) with identical injected
    noise - PSNR of the restored RGB (pred_xstart[:, :3], data range 2.0) per image, and phi.  Free-running chains separate
    through clamp sign flips (the reference against its own fp64 twin reaches 55 dB after 250 steps)."""
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import FusedStepper
    a = arguments_from_file(os.path.join(ROOT, "configs", "osmosis_simulation_sample_config.yaml"))
    B, T = 2, 12
    opc0 = dict(a.measurement["operator"])
    ph = lambda k: [float(v) for v in str(opc0[k]).split(",")]
    y = torch.cat([synth_measurement(20 + i, S, ph("phi_ab"), ph("phi_ab"), ph("phi_inf"), depth_type=opc0.get("depth_type"))[0]
                   for i in range(B)]).to(DEV)
    g = torch.Generator().manual_seed(2)
    x_T = torch.randn(1, 4, S, S, generator=g).repeat(B, 1, 1, 1).to(DEV)
    noises = [torch.randn(1, 4, S, S, generator=g).repeat(B, 1, 1, 1).to(DEV) for _ in range(T)]
    res = {}
    for mode in ("fp32", "tc"):
        m = full_model(mode)
        opc = dict(opc0); opc["batch_size"] = B
        op = get_operator(device=DEV, **opc)
        cond = get_conditioning_method(a.conditioning["method"], op, get_noise(**a.measurement["noise"]), **a.conditioning["params"],
                                       **a.sample_pattern, **a.aux_loss)
        d = dict(a.diffusion); d["timestep_respacing"] = T
        sampler = create_sampler(**d)
        img = x_T.clone()
        stepper = FusedStepper(sampler, m, cond, img, y, a.sample_pattern, cuda_graph=(mode == "tc"))
        for k, idx in enumerate(range(T)[::-1]):
            stepper._draw_into = lambda buf, _k=k: buf.copy_(noises[_k]) if buf.shape[1] == 4 else buf.zero_()
            stepper.step(idx)
        torch.cuda.synchronize()
        res[mode] = (stepper.st["x0"].clone(), op.phi.clone(), stepper.st["losses"].clone())
        assert torch.isfinite(res[mode][0]).all()
    mse = ((res["tc"][0][:, :3] - res["fp32"][0][:, :3]) ** 2).mean(dim=(1, 2, 3))
    psnr = 10 * torch.log10(4.0 / mse)
    print("full-size 12-step chain PSNR (dB) product vs exact mode:", [round(float(v), 1) for v in psnr])
    assert float(psnr.min()) > 50.0, psnr        # measured 86 dB on B200
    assert float((res["tc"][1] - res["fp32"][1]).abs().max()) < 1e-3
    assert rel_err(res["tc"][2][:, 0].cpu(), res["fp32"][2][:, 0].cpu()) < 5e-2
