"""CPU: the oracle's input-pipeline restatement against golden vectors produced by the unmodified reference datasets +
the transform chain of osmosis_sampling.py:46-49 under torchvision (tests/golden/make_golden_pre.py -> pre_golden.npz).
The restatement is BIT-EXACT for every case but the 2.8x down-scale, where ATen's vectorised loop sums the 7-tap
filters in a different order: tolerance there 5e-7 absolute on values in [-1, 1] (2 ulp at 1.0 after the Normalize)."""
import os

import numpy as np
import pytest

from oracle import osmosis_oracle as orc
from tests.golden.cases import PRE_CASES, pre_inputs, pre_check

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "pre_golden.npz"))
ATOL = 5e-7
EXACT = [n for n in PRE_CASES if n != "big720x1280"]


@pytest.mark.parametrize("name", list(PRE_CASES))
def test_preprocess_matches_torchvision_chain(name):
    got = orc.preprocess_image(pre_inputs(name))
    assert got.shape == (3, 256, 256) and got.dtype == np.float32
    assert pre_check(GOLD, name, got, 0.0 if name in EXACT else ATOL)


@pytest.mark.parametrize("name", ["land300x400", "up200x320"])
def test_degamma_matches_reference(name):
    got = orc.preprocess_image(pre_inputs(name), degamma=True)
    assert pre_check(GOLD, name + ":degamma", got, 5e-7)     # + 1 ulp of pow


def test_identity_size_is_exact_and_crop_is_centred():
    a = pre_inputs("same256")
    got = orc.preprocess_image(a)
    want = (a.astype(np.float32) / np.float32(255) - np.float32(0.5)) / np.float32(0.5)
    assert np.array_equal(got, want.transpose(2, 0, 1))
    # a wide image of constant columns: the crop keeps the centre columns (round-half-to-even offset)
    w = np.zeros((256, 341, 3), np.uint8); w[:, 42:42 + 256] = 255
    assert np.array_equal(orc.preprocess_image(w), np.ones((3, 256, 256), np.float32))
