"""CPU: host side of the input pipeline - natural ordering, dataset decoding, rank sharding (no device work)."""
import numpy as np
import pytest

from osmosis_diffusion_code_b200.osmosis_utils import data as D


def test_natsorted_orders_numbers_numerically():
    assert D.natsorted(["b10.png", "b2.png", "a.png", "b1.png", "B3.png"]) == ["B3.png", "a.png", "b1.png", "b2.png", "b10.png"]


def test_datasets_decode_like_the_reference(tmp_path):
    from PIL import Image
    rs = np.random.RandomState(0)
    for sub in ("in", "rgb", "depth"):
        (tmp_path / sub).mkdir()
    a = rs.randint(0, 256, (40, 50, 3)).astype(np.uint8)
    d16 = rs.randint(0, 65536, (40, 50)).astype(np.uint16)
    Image.fromarray(a).save(tmp_path / "in" / "x1.png")
    Image.fromarray(a[::-1].copy()).save(tmp_path / "rgb" / "x1.png")
    Image.fromarray(d16).save(tmp_path / "depth" / "x1.png")
    ds = D.ImagesFolder(str(tmp_path / "in"))
    img, name = ds[0]
    assert name == "x1.png" and np.array_equal(img, a)
    gt = D.ImagesFolder_GT(str(tmp_path / "in"), str(tmp_path / "rgb"), str(tmp_path / "depth"))
    (y, rgb, depth), name = gt[0]
    assert np.array_equal(y, a) and np.array_equal(rgb, a[::-1]) and np.array_equal(depth, (d16 // 256).astype(np.uint8))   # data.py:97-99


def test_shard_indices_partition_the_dataset():
    class DS:
        def __len__(self): return 11
    parts = [D.ShardedImageLoader(DS(), 2, rank=r, world=4).indices for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(11))
    assert D.ShardedImageLoader(DS(), 2, rank=0, world=4, stop_after=5).indices == [0, 4]
    assert len(D.ShardedImageLoader(DS(), 2, rank=0, world=4)) == 2


def test_preprocess_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(Exception):
        D.preprocess_batch([np.zeros((8, 8, 3), np.uint8)])
