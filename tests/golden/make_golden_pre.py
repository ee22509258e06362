"""Golden vectors of the input pipeline: the transform chain the reference builds in osmosis_sampling.py:46-49

    transforms.Compose([ToTensor(), Resize(size=256), CenterCrop(size=[256, 256]), Normalize((0.5,)*3, (0.5,)*3)])

(restated here verbatim because it is a local variable of the reference's `main`), applied by the UNMODIFIED reference
datasets `osmosis_utils/data.py` (`ImagesFolder`, `ImagesFolder_GT`) to seeded PNG files, plus the de-gamma of
osmosis_sampling.py:170-175.  torchvision 0.26 / torch 2.11 CPU kernels.

Run in the build container only (needs /root/reference, PIL, cv2 is stubbed unless present):
    python tests/golden/make_golden_pre.py
Only outputs are stored (tests/golden/pre_golden.npz); the inputs are regenerated from the seeds by pre_inputs().
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for name in ("natsort", "cv2", "matplotlib", "matplotlib.pyplot"):   # imported by the reference, unused on this path
    try:
        __import__(name)
    except ImportError:
        sys.modules[name] = types.ModuleType(name)
if not hasattr(sys.modules["natsort"], "natsorted"):
    sys.modules["natsort"].natsorted = sorted
if not hasattr(sys.modules["matplotlib"], "pyplot"):
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
import osmosis_utils.data as refdata  # noqa: E402  (the reference)

from tests.golden.cases import PRE_CASES, pre_inputs, FULL_CASE  # noqa: E402


def main():
    transform = transforms.Compose([transforms.ToTensor(),
                                    transforms.Resize(size=256),
                                    transforms.CenterCrop(size=[256, 256]),
                                    transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    out = {}
    with tempfile.TemporaryDirectory() as d:
        names = [n for n in PRE_CASES if PRE_CASES[n][3] != "grey"]
        for n in names:
            Image.fromarray(pre_inputs(n)).save(os.path.join(d, n + ".png"))
        ds = refdata.ImagesFolder(d, transform)
        for i in range(len(ds)):
            img, fname = ds[i]
            n = os.path.splitext(fname)[0]
            out[n] = img.numpy()
            y_n_tmp = 0.5 * (img + 1)                                   # osmosis_sampling.py:173-175
            out[n + ":degamma"] = (2 * torch.pow(y_n_tmp, 2.2) - 1).numpy()
    # the ground-truth depth path of ImagesFolder_GT (data.py:94-107): 8-bit grey -> .convert("RGB") -> same transform
    for n in PRE_CASES:
        if PRE_CASES[n][3] == "grey":
            out[n] = transform(Image.fromarray(pre_inputs(n)).convert(mode="RGB")).numpy()
    # keep the fixture small: one case in full, the others as every 3rd pixel plus a float64 checksum of the whole plane
    small = {}
    for k, v in out.items():
        if k == FULL_CASE:
            small[k] = v
        else:
            small[k + "|sub3"] = np.ascontiguousarray(v[:, ::3, ::3])
            small[k + "|sum"] = np.array([v.astype(np.float64).sum(), np.abs(v.astype(np.float64)).sum()])
    out = small
    np.savez_compressed(os.path.join(HERE, "pre_golden.npz"), **out)
    print("wrote", len(out), "arrays", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
