"""Golden vectors of the post-processing helpers, produced by the UNMODIFIED reference functions
(osmosis_utils/utils.py: min_max_norm_range, min_max_norm_range_percentile, convert_depth) on seeded inputs.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_post.py
Only outputs are stored (tests/golden/post_golden.npz); the inputs are regenerated from the seeds by post_inputs().
"""
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
import osmosis_utils.utils as ref  # noqa: E402  (the reference)

from tests.golden.cases import post_inputs, POST_CASES  # noqa: E402


def main():
    out = {}
    for name in POST_CASES:
        d = post_inputs(name)
        out[f"{name}:mm"] = ref.min_max_norm_range(d).numpy()
        out[f"{name}:pmm_03_99"] = ref.min_max_norm_range_percentile(d, vmin=0, vmax=1, percent_low=0.03, percent_high=0.99).numpy()
        out[f"{name}:pmm_05_99"] = ref.min_max_norm_range_percentile(d, percent_low=0.05, percent_high=0.99).numpy()
        out[f"{name}:pmm_range"] = ref.min_max_norm_range_percentile(d, vmin=-1, vmax=2, percent_low=0.25, percent_high=0.5).numpy()
        out[f"{name}:gamma"] = ref.convert_depth(d.repeat(3, 1, 1), depth_type="gamma", value="1.4,1.4,1").numpy()
    np.savez_compressed(os.path.join(HERE, "post_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
