"""CPU: the oracle's restatement of BASELINE config 1 (`GaussianDiffusion.inverse`, osmosis_utils/diffusion.py:59-130) against
the unmodified reference's own run (tests/golden/make_golden_uncond.py -> uncond_golden.npz), RNG drawn in its order."""
import os

import numpy as np
import torch

from oracle import osmosis_oracle as orc
from tests.golden.cases import UNCOND_CASE, SMALL_HW
from tests.helpers import small_state_dict, small_cfg, maxdiff

GOLD = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "uncond_golden.npz")))


def test_uncond_inverse_matches_reference():
    c = UNCOND_CASE
    torch.manual_seed(c["seed"])
    x_T = torch.randn(1, 4, SMALL_HW, SMALL_HW)                      # diffusion.py:75
    x, x0 = orc.uncond_inverse(small_state_dict(), small_cfg(), c["T"], "linear", x_T, lambda t: torch.randn(1, 4, SMALL_HW, SMALL_HW),
                               start_t=c["start_t"], steps=c["steps"])
    assert maxdiff(x, GOLD["x"]) < 2e-5
    rgb = torch.clamp(0.5 * (x0[0, :3] + 1), 0, 1)
    assert maxdiff(rgb, GOLD["x_start_rgb"]) < 2e-5
    d01 = (0.5 * (x0[0, 3] + 1)).unsqueeze(0)
    assert maxdiff(orc.min_max_norm_range_percentile(d01, 0.0, 1.0, 0.05, 0.99), GOLD["x_depth_pmm"]) < 1e-4
