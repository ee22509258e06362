"""GPU: the device post-processing kernels (through the reference-named host functions and the C ABI) against the golden
vectors of the unmodified reference helpers and the oracle.  Bit-exact for the min-max / percentile normalisation and the
colour lookup (order statistics + fp32 elementwise with torch's rounding); 2e-6 for the exp-based image formation."""
import os

import numpy as np
import pytest
import torch

from oracle import osmosis_oracle as orc
from osmosis_diffusion_code_b200.osmosis_utils import utils as U
from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator
from tests.golden.cases import POST_CASES, post_inputs, CASES, case_inputs
from tests.helpers import load_yaml_cfg, oracle_specs_from_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_golden.npz"))


@pytest.mark.parametrize("name", list(POST_CASES))
def test_minmax_and_percentile_bit_exact_vs_reference_golden(name):
    d = post_inputs(name).to(DEV)
    assert np.array_equal(U.min_max_norm_range(d).cpu().numpy(), GOLD[f"{name}:mm"])
    assert np.array_equal(U.min_max_norm_range_percentile(d, 0, 1, 0.03, 0.99).cpu().numpy(), GOLD[f"{name}:pmm_03_99"])
    assert np.array_equal(U.min_max_norm_range_percentile(d, percent_low=0.05, percent_high=0.99).cpu().numpy(), GOLD[f"{name}:pmm_05_99"])
    assert np.array_equal(U.min_max_norm_range_percentile(d, -1, 2, 0.25, 0.5).cpu().numpy(), GOLD[f"{name}:pmm_range"])


def test_percentile_full_size_batch_vs_torch_quantile():
    """256x256 planes, batch of 5 (per-image semantics): bit-exact against the oracle (torch.quantile on the CPU)."""
    g = torch.Generator().manual_seed(21)
    d = torch.randn(5, 1, 256, 256, generator=g) * torch.tensor([0.1, 1.0, 3.0, 10.0, 1e-3]).view(5, 1, 1, 1)
    got = U.min_max_norm_range_percentile(d.to(DEV), 0, 1, 0.03, 0.99).cpu()
    for b in range(5):
        assert torch.equal(got[b], orc.min_max_norm_range_percentile(d[b], 0, 1, 0.03, 0.99)), b
    got = U.min_max_norm_range(d.to(DEV)).cpu()
    for b in range(5):
        assert torch.equal(got[b], orc.min_max_norm_range(d[b])), b


def test_colormap_lookup():
    lut = np.random.RandomState(0).rand(256, 3).astype(np.float32)
    g = torch.Generator().manual_seed(2)
    img = torch.rand(3, 1, 40, 24, generator=g)
    img[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 0.5, 255.0 / 256])
    got = U.depth_tensor_to_color_image(img.to(DEV), lut=lut).cpu()
    for b in range(3):
        assert torch.equal(got[b], orc.apply_colormap(img[b, 0], lut))
    one = U.depth_tensor_to_color_image(img[:1].to(DEV), lut=lut).cpu()   # reference shapes: [1,1,H,W] -> [3,H,W]
    assert one.shape == (3, 40, 24) and torch.equal(one, got[0])
    tab = U.colormap_table("viridis")
    assert tab.shape == (256, 3) and abs(float(tab[0, 2]) - 0.33) < 0.02 and abs(float(tab[255, 0]) - 0.99) < 0.02


@pytest.mark.parametrize("cname", list(CASES))
def test_postprocess_samples_vs_oracle(cname):
    c = CASES[cname]
    cfg = load_yaml_cfg(c["yaml"], c["respacing"])
    B = 3
    opcfg = dict(cfg["measurement"]["operator"]); opcfg["batch_size"] = B
    op = get_operator(device=DEV, **opcfg)
    tab, ospec, gspec, phis, names = oracle_specs_from_cfg(cfg, B)
    g = torch.Generator().manual_seed(31)
    y1, xgt = case_inputs("meas:" + cname)
    H = y1.shape[-1]
    x0 = (xgt + 0.4 * torch.randn(B, 4, H, H, generator=g)).contiguous()
    y = (y1 + 0.05 * torch.randn(B, 3, H, H, generator=g)).contiguous()
    # distinct phi per image
    for i, n in enumerate(names):
        p = getattr(op, n)
        p.mul_(1.0 + 0.1 * torch.arange(B, device=DEV).view(B, 1, 1, 1))
        phis[i] = p.detach().cpu().clone()
    r = U.postprocess_samples(op, x0.to(DEV), y.to(DEV))
    torch.cuda.synchronize()
    for b in range(B):
        o = orc.postprocess(ospec, x0[b:b + 1], y[b:b + 1], [p[b:b + 1] for p in phis])
        assert torch.equal(r["sample_rgb_01_clip"][b].cpu(), o["sample_rgb_01_clip"])
        assert torch.equal(r["sample_depth_mm"][b].cpu(), o["sample_depth_mm"])
        assert torch.equal(r["sample_depth_vis_pmm"][b].cpu(), o["sample_depth_vis_pmm"])
        for k in ("degraded_image", "sample_rgb_recon"):
            s = float(o[k].abs().max())
            assert float((r[k][b].cpu() - o[k]).abs().max()) < 2e-6 * max(1.0, s), k
        assert abs(float(r["norm_loss"][b]) - float(o["norm_loss"])) < 1e-5 * max(1.0, float(o["norm_loss"]))
    assert r["sample_depth_vis_pmm_color"].shape == (B, 3, H, H)


def test_p_sample_loop_record_builds_the_process_grid(tmp_path):
    """`record=True` (gaussian_diffusion.py:310-333): frames at idx % record_every == 0 and idx == 0, grid saved as PNG."""
    from tests.test_path_gpu import _native_objects, model
    cname = "osmosis"
    cfg, op, cond, sampler = _native_objects(cname, 1)
    y, _ = case_inputs("meas:" + cname)
    torch.manual_seed(0)
    x_start = torch.randn(1, 4, *y.shape[2:], device=DEV)
    sampler.p_sample_loop(model=model("fp32"), x_start=x_start, measurement=y.to(DEV), measurement_cond_fn=cond.conditioning,
                          record=True, save_root=None, pretrain_model="osmosis", rgb_guidance=False,
                          sample_pattern=cfg["sample_pattern"], record_every=2, save_grids_path=str(tmp_path),
                          original_file_name="img7")
    rec = sampler.last_record
    n = len([i for i in range(sampler.num_timesteps) if i % 2 == 0 or i == 0])
    assert rec["rgb"].shape[0] == n and rec["depth_color"].shape == rec["rgb"].shape
    assert float(rec["rgb"].min()) >= 0 and float(rec["rgb"].max()) <= 1
    assert os.path.exists(os.path.join(str(tmp_path), "img7_process.png"))
    assert rec["grid"].shape[0] == 3
