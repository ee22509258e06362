"""GPU: the device input pipeline (csrc/preprocess.cu through osm_preprocess_image) against the oracle restatement -
bit-exact, the kernels mirror its rounding op by op - and against the golden vectors the unmodified reference datasets +
torchvision chain produced (tests/golden/make_golden_pre.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import osmosis_oracle as orc
from tests.golden.cases import PRE_CASES, pre_inputs, pre_check

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "pre_golden.npz"))


def test_batch_of_mixed_sizes_matches_oracle_bit_exactly_and_reference_golden():
    from osmosis_diffusion_code_b200.osmosis_utils.data import preprocess_batch
    names = list(PRE_CASES)
    imgs = [pre_inputs(n) for n in names]
    out = preprocess_batch(imgs).cpu().numpy()
    assert out.shape == (len(names), 3, 256, 256)
    for k, n in enumerate(names):
        assert np.array_equal(out[k], orc.preprocess_image(imgs[k])), n
        assert pre_check(GOLD, n, out[k], 0.0 if n != "big720x1280" else 5e-7), n


def test_empty_and_extreme_inputs():
    """An empty batch (a rank with no images left), a 1-pixel-high strip, a constant image and the saturated values."""
    from osmosis_diffusion_code_b200.osmosis_utils.data import preprocess_batch
    assert tuple(preprocess_batch([]).shape) == (0, 3, 256, 256)
    strip = np.full((1, 700, 3), 200, np.uint8)                     # up-scaled 256x vertically, cropped horizontally
    const = np.full((300, 300), 0, np.uint8)                        # grey, all black -> -1
    white = np.full((256, 256, 3), 255, np.uint8)                   # identity size, all white -> +1
    out = preprocess_batch([strip, const, white]).cpu().numpy()
    for k, img in enumerate((strip, const, white)):
        assert np.array_equal(out[k], orc.preprocess_image(img)), k
    assert np.all(out[1] == -1.0) and np.all(out[2] == 1.0)
    assert np.allclose(out[0], 2 * np.float32(200) / 255 - 1, atol=3e-7)
    big = np.random.RandomState(1).randint(0, 256, (1500, 2100, 3)).astype(np.uint8)     # 5.9x down-scale: 13-tap filters
    assert np.array_equal(preprocess_batch([big]).cpu().numpy()[0], orc.preprocess_image(big))


def test_degamma_fused_and_standalone():
    from osmosis_diffusion_code_b200.osmosis_utils.data import preprocess_batch, degamma_input
    imgs = [pre_inputs("land300x400"), pre_inputs("up200x320")]
    fused = preprocess_batch(imgs, degamma=True)
    alone = degamma_input(preprocess_batch(imgs))
    assert torch.equal(fused, alone)
    for k, n in enumerate(["land300x400", "up200x320"]):
        assert pre_check(GOLD, n + ":degamma", fused[k].cpu().numpy(), 1e-6), n     # powf vs torch.pow: <= 2 ulp


def test_sharded_loader_covers_the_dataset_once(tmp_path):
    from PIL import Image
    from osmosis_diffusion_code_b200.osmosis_utils.data import ImagesFolder, ShardedImageLoader
    rs = np.random.RandomState(3)
    raw = {}
    for i in (1, 2, 10, 11, 3):
        a = rs.randint(0, 256, (260 + i, 300, 3)).astype(np.uint8)
        Image.fromarray(a).save(tmp_path / f"im{i}.png")
        raw[f"im{i}.png"] = a
    ds = ImagesFolder(str(tmp_path))
    assert ds.images_list == ["im1.png", "im2.png", "im3.png", "im10.png", "im11.png"]
    seen = {}
    for rank in range(2):
        for y, names, extras in ShardedImageLoader(ds, batch_per_rank=2, rank=rank, world=2):
            assert y.shape[1:] == (3, 256, 256) and y.is_cuda and extras == {}
            for k, n in enumerate(names):
                assert n not in seen
                seen[n] = y[k].cpu().numpy()
    assert sorted(seen) == sorted(raw)
    for n, a in raw.items():
        assert np.array_equal(seen[n], orc.preprocess_image(a))
