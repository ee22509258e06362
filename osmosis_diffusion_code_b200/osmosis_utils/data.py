"""Input pipeline: the reference's datasets (`osmosis_utils/data.py:15-109`) and its transform chain
(`osmosis_sampling.py:46-49`: ToTensor -> Resize(256) -> CenterCrop -> Normalize, plus the de-gamma of :170-175), with the
pixel work on the device.

The reference decodes with PIL and transforms one image at a time on the CPU inside a `DataLoader` (`:117`).  Here the
datasets decode on the host (file decoding is not GPU work) and hand out the raw uint8 pixels; `preprocess_batch` stages a
batch of them in pinned memory, copies it to the device asynchronously and runs `osm_preprocess_image` per image on the
current stream, writing straight into the `[B,3,256,256]` measurement tensor the sampler consumes.  `ShardedImageLoader`
yields such batches for one rank of a batch-sharded run (rank r of R takes images r, r+R, ... of every global batch).

No CPU fallback: `preprocess_batch` raises without CUDA / the native library.
"""
from __future__ import annotations

import glob
import os
import re
from os.path import join as pjoin

import numpy as np
import torch

from .. import lib as _lib

IMAGE_SIZE = 256


def natsorted(names):
    """Natural ordering of file names ("img2" < "img10"), what the reference gets from the `natsort` package."""
    def key(s):
        return [(0, int(p)) if p.isdigit() else (1, p) for p in re.split(r"(\d+)", os.fspath(s)) if p != ""]
    return sorted(names, key=key)


def _decode(path) -> np.ndarray:
    """Decoded pixels as uint8 [H,W,3] (grey images stay [H,W]); 16-bit depth maps are reduced to 8 bits like data.py:97-99."""
    from PIL import Image
    img = Image.open(path)
    if img.mode in ("I;16", "I;16B", "I"):
        return (np.asarray(img).astype(np.int64) // 256).astype(np.uint8)
    if img.mode not in ("RGB", "L"):
        img = img.convert("RGB")
    return np.ascontiguousarray(np.asarray(img))


class ImagesFolder:
    """data.py:15-37: every file of `root_dir` in natural order -> (image, file name).  With `transform=None` the image is
    the raw uint8 array (feed it to `preprocess_batch`); a callable transform is applied like the reference does."""

    def __init__(self, root_dir, transform=None):
        self.root_dir = root_dir
        self.images_list = natsorted(os.listdir(root_dir))
        self.transform = transform

    def __len__(self):
        return len(self.images_list)

    def __getitem__(self, idx):
        image = _decode(os.path.join(self.root_dir, self.images_list[idx]))
        if self.transform is not None:
            image = self.transform(image)
        return image, self.images_list[idx]


class ImagesFolder_GT:
    """data.py:73-109: degraded image + ground-truth RGB + ground-truth depth (grey, replicated to 3 channels)."""

    def __init__(self, root_dir, gt_rgb_dir, gt_depth_dir, transform=None):
        self.gt_rgb_list = natsorted(glob.glob(pjoin(gt_rgb_dir, "*.*")))
        self.gt_depth_list = natsorted(glob.glob(pjoin(gt_depth_dir, "*.*")))
        self.images_list = natsorted(glob.glob(pjoin(root_dir, "*.*")))
        self.transform = transform

    def __len__(self):
        return len(self.gt_rgb_list)

    def __getitem__(self, idx):
        name = os.path.basename(self.images_list[idx])
        items = [_decode(self.images_list[idx]), _decode(self.gt_rgb_list[idx]), _decode(self.gt_depth_list[idx])]
        if items[2].ndim == 3:
            items[2] = np.ascontiguousarray(items[2][:, :, 0])
        if self.transform is not None:
            items = [self.transform(i) for i in items]
        return items, name


def preprocess_batch(images, device=None, size: int = IMAGE_SIZE, degamma: bool = False, out: torch.Tensor | None = None):
    """uint8 images (list of [H,W,3] / [H,W] numpy arrays or PIL images, any sizes) -> float32 [B,3,size,size] in [-1,1] on
    the device: the reference's transform chain (osmosis_sampling.py:46-49) and, with `degamma`, :170-175.

    One pinned staging buffer and one device buffer hold the whole batch; the copy is asynchronous and the kernels are
    queued behind it on the current stream, so the call returns without synchronising."""
    _lib.require_cuda()
    L = _lib.load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    arrs = []
    for im in images:
        a = np.asarray(im)
        if a.dtype != np.uint8 or a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] != 3):
            raise ValueError(f"expected uint8 [H,W,3] or [H,W] images, got {a.dtype} {a.shape}")
        arrs.append(np.ascontiguousarray(a))
    B = len(arrs)
    if B == 0:   # a rank whose shard of the last global batch is empty
        return out if out is not None else torch.empty(0, 3, size, size, dtype=torch.float32, device=dev)
    offs = np.cumsum([0] + [(a.size + 255) // 256 * 256 for a in arrs])
    stage = torch.empty(int(offs[-1]), dtype=torch.uint8).pin_memory()
    sv = stage.numpy()
    for a, o in zip(arrs, offs):
        sv[o:o + a.size] = a.reshape(-1)
    with torch.cuda.device(dev):
        raw = stage.to(dev, non_blocking=True)
        if out is None:
            out = torch.empty(B, 3, size, size, dtype=torch.float32, device=dev)
        scratch = torch.empty(3 * max(a.shape[0] for a in arrs) * size, dtype=torch.float32, device=dev)
        for b, (a, o) in enumerate(zip(arrs, offs)):
            H, W = a.shape[:2]
            ch = 1 if a.ndim == 2 else 3
            _lib.check(L.osm_preprocess_image(_lib.C.c_void_p(raw.data_ptr() + int(o)), H, W, ch, W * ch, _lib.ptr(scratch),
                                              _lib.C.c_void_p(out.data_ptr() + b * 3 * size * size * 4), size, int(bool(degamma)),
                                              _lib.stream()))
        raw.record_stream(torch.cuda.current_stream())
    return out


def degamma_input(y: torch.Tensor) -> torch.Tensor:
    """osmosis_sampling.py:173-175 on a device tensor: 2 (0.5 (y + 1))^2.2 - 1."""
    _lib.require_cuda()
    y = y.contiguous().float()
    if not y.is_cuda:
        raise _lib.OsmError("degamma_input runs on CUDA tensors only (no CPU fallback)")
    out = torch.empty_like(y)
    _lib.check(_lib.load().osm_degamma(_lib.ptr(y), _lib.ptr(out), y.numel(), _lib.stream()))
    return out


class ShardedImageLoader:
    """Replaces the reference's sequential `DataLoader(dataset, batch_size, shuffle=False)` loop (osmosis_sampling.py:55-61,
    :117) for a batch-sharded run: every iteration yields `(measurement [b,3,256,256] on the device, names, extras)` for
    THIS rank's slice of the next global batch.  Image i of the dataset goes to rank (i mod world) - images are
    independent, so no collective is involved.  `extras` holds the ground-truth tensors for a GT dataset."""

    def __init__(self, dataset, batch_per_rank: int, rank: int = 0, world: int = 1, device=None, degamma: bool = False,
                 stop_after: int = -1, size: int = IMAGE_SIZE):
        self.ds, self.b, self.rank, self.world, self.device, self.degamma = dataset, batch_per_rank, rank, world, device, degamma
        self.size = size
        n = len(dataset) if stop_after is None or stop_after < 0 else min(len(dataset), stop_after)
        from ..sharding import shard_indices
        self.indices = shard_indices(n, rank, world)

    def __len__(self):
        return (len(self.indices) + self.b - 1) // self.b

    def __iter__(self):
        for s in range(0, len(self.indices), self.b):
            items = [self.ds[i] for i in self.indices[s:s + self.b]]
            names = [n for _, n in items]
            if isinstance(items[0][0], (list, tuple)):
                y = preprocess_batch([it[0][0] for it in items], self.device, size=self.size, degamma=self.degamma)
                gt_rgb = preprocess_batch([it[0][1] for it in items], self.device, size=self.size)
                gt_depth = preprocess_batch([it[0][2] for it in items], self.device, size=self.size)
                yield y, names, dict(gt_rgb=gt_rgb, gt_depth=gt_depth)
            else:
                yield preprocess_batch([it[0] for it in items], self.device, size=self.size, degamma=self.degamma), names, {}
