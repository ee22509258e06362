"""CPU: the oracle's post-processing restatement against golden vectors produced by the unmodified reference helpers
(tests/golden/make_golden_post.py -> post_golden.npz).  Bit-exact: these are elementwise fp32 / order-statistic ops."""
import os

import numpy as np
import pytest
import torch

from oracle import osmosis_oracle as orc
from tests.golden.cases import POST_CASES, post_inputs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_golden.npz"))


@pytest.mark.parametrize("name", list(POST_CASES))
def test_minmax_and_percentile_match_reference(name):
    d = post_inputs(name)
    assert np.array_equal(orc.min_max_norm_range(d).numpy(), GOLD[f"{name}:mm"])
    assert np.array_equal(orc.min_max_norm_range_percentile(d, 0, 1, 0.03, 0.99).numpy(), GOLD[f"{name}:pmm_03_99"])
    assert np.array_equal(orc.min_max_norm_range_percentile(d, 0, 1, 0.05, 0.99).numpy(), GOLD[f"{name}:pmm_05_99"])
    assert np.array_equal(orc.min_max_norm_range_percentile(d, -1, 2, 0.25, 0.5).numpy(), GOLD[f"{name}:pmm_range"])
    assert np.array_equal(orc.convert_depth(d.repeat(3, 1, 1), "gamma", (1.4, 1.4, 1.0)).numpy(), GOLD[f"{name}:gamma"])


def test_colormap_lookup_semantics():
    lut = np.stack([np.arange(256), 255 - np.arange(256), np.arange(256) % 7], axis=1).astype(np.float32) / 255.0
    img = torch.tensor([[0.0, 1.0 / 256, 0.5, 0.999, 1.0, 0.25 - 1e-7]])
    out = orc.apply_colormap(img, lut)
    assert out.shape == (3, 1, 6)
    assert [int(round(float(v) * 255)) for v in out[0, 0]] == [0, 1, 128, 255, 255, 63]


def test_postprocess_is_consistent_with_the_operator():
    """degraded_image == 2 A_phi(x0) - 1 with the (golden-pinned) operator forward; the restored image inverts the model:
    applying the operator to (recon, depth) gives back the measurement."""
    spec = orc.OperatorSpec("underwater_physical_revised", "gamma", (1.4, 1.4, 1.0), (1e-5, 1e-5, 1e-5))
    g = torch.Generator().manual_seed(5)
    x0 = torch.rand(1, 4, 24, 24, generator=g) * 2 - 1
    phis = [torch.tensor([1.1, 0.95, 0.9]).view(1, 3, 1, 1), torch.tensor([0.9, 0.8, 0.7]).view(1, 3, 1, 1),
            torch.tensor([0.14, 0.29, 0.49]).view(1, 3, 1, 1)]
    y = 2 * orc.operator_forward(spec, x0, phis) - 1
    r = orc.postprocess(spec, x0, y, phis)
    assert float((r["degraded_image"] - y[0]).abs().max()) < 1e-6 and float(r["norm_loss"]) < 1e-4
    x_rec = torch.cat([2 * r["sample_rgb_recon"] - 1, x0[0, 3:4]], 0)[None]
    assert float((2 * orc.operator_forward(spec, x_rec, phis) - 1 - y).abs().max()) < 1e-5
    assert float(r["sample_rgb_01_clip"].min()) >= 0 and float(r["sample_rgb_01_clip"].max()) <= 1
