"""Golden vectors of BASELINE config 1 (`RGBD_prior_sampling.py`): the UNMODIFIED reference's
`osmosis_utils/diffusion.py` `GaussianDiffusion(T=50, 'linear').inverse(...)` on the small UNet, the last 6 steps of the
chain (start_t = 6: the record at t == 1 is what defines the returned x_start_rgb / x_depth; without it the reference raises
UnboundLocalError at :130).  matplotlib is not installed here, so `plt.get_cmap` is stood in by an identity "colour map"
(value -> (v, v, v)): the stored depth is therefore the percentile-normalised map BEFORE colouring.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_uncond.py
"""
import os
import sys
import tempfile
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
sys.modules["matplotlib.pyplot"].get_cmap = lambda name: (lambda a: np.stack([a, a, a, np.ones_like(a)], axis=-1))
sys.path.insert(0, "/root/reference")

from guided_diffusion.unet import create_model  # noqa: E402  (the reference)
from osmosis_utils.diffusion import GaussianDiffusion  # noqa: E402

from osmosis_diffusion_code_b200.synthetic import synth_state_dict  # noqa: E402
from tests.golden.cases import SMALL_UNET, SMALL_HW, UNCOND_CASE  # noqa: E402

torch.set_num_threads(8)


def main():
    m = create_model(**SMALL_UNET, model_path="/nonexistent")
    specs = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    m.load_state_dict(synth_state_dict(specs, SMALL_UNET["num_channels"], seed=7, delta=0.05), strict=True)
    m.eval()
    c = UNCOND_CASE
    diffusion = GaussianDiffusion(T=c["T"], schedule="linear")
    torch.manual_seed(c["seed"])
    with tempfile.TemporaryDirectory() as d:
        x, (rgb, depth) = diffusion.inverse(net=m, shape=(4, SMALL_HW, SMALL_HW), image_channels=4, steps=c["steps"], start_t=c["start_t"],
                                            device="cpu", record_process=True, record_every=200, save_path=d, image_idx=0)
        assert os.path.exists(os.path.join(d, "image_0_process.png"))
    out = {"x": x.numpy(), "x_start_rgb": rgb.numpy(), "x_depth_pmm": depth.numpy()[0:1]}
    np.savez_compressed(os.path.join(HERE, "uncond_golden.npz"), **out)
    print("wrote", {k: v.shape for k, v in out.items()}, float(np.abs(out["x"]).max()))


if __name__ == "__main__":
    main()
