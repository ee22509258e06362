"""End-to-end sampling run on the device: the body of the reference's `osmosis_sampling.py:main` (:36-330) with every
pixel-touching stage native - decode -> device input pipeline -> guided sampling -> device post-processing -> images.

    python -m osmosis_diffusion_code_b200.sampling -c configs/osmosis_sample_config.yaml [-d 0] [--batch 8]
    torchrun --nproc-per-node 8 -m osmosis_diffusion_code_b200.sampling -c ... --batch 32      # one process per GPU

What differs from the reference's driver, on purpose:
  * images are processed `--batch` at a time per GPU instead of one by one (a batch is B independent reference runs:
    per-image loss norm, per-image phi, shared x_T / step noise exactly as the reference's per-image reseed gives);
  * with several processes, image i of the dataset goes to rank i mod world (no collective: images are independent);
  * the per-image CPU post-processing block (:207-292) is `postprocess_samples` on the device; only the finished 8-bit
    images cross to the host.
Logging / dated output directories / the configuration dump of the reference's CLI are not reproduced.
"""
from __future__ import annotations

import argparse
import os
from os.path import join as pjoin

import numpy as np
import torch

from .guided_diffusion.condition_methods import get_conditioning_method
from .guided_diffusion.gaussian_diffusion import create_sampler
from .guided_diffusion.measurements import get_noise, get_operator
from .guided_diffusion.unet import create_model
from .osmosis_utils import data as datao
from .osmosis_utils import utils as utilso


def _save_png(t: torch.Tensor, path: str):
    """[3,H,W] or [1,H,W] in [0,1] -> 8-bit PNG (what torchvision's to_pil_image does: mul(255).byte())."""
    from PIL import Image
    a = t.detach().clamp(0, 1).mul(255).byte().cpu().numpy()
    Image.fromarray(a[0] if a.shape[0] == 1 else a.transpose(1, 2, 0)).save(path)


def run_sampling(args, device=None, batch_per_rank: int = 1, rank: int = 0, world: int = 1, out_dir: str | None = None, model=None,
                 image_size: int = datao.IMAGE_SIZE, max_steps: int | None = None, cuda_graph: bool = True):
    """Runs the config `args` (from `arguments_from_file`) over its dataset.  Returns a list of per-image dicts
    {name, loss, norm_loss, phi_*} and, when `out_dir` is given and `args.save_singles`, writes
    <out_dir>/single_images/{input,rgb,depth_color,depth_raw}/<name>.png like the reference (:300-320)."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    dc = args.data
    if dc.get("ground_truth"):
        dataset = datao.ImagesFolder_GT(root_dir=dc["root"], gt_rgb_dir=dc["gt_rgb"], gt_depth_dir=dc["gt_depth"])
    else:
        dataset = datao.ImagesFolder(dc["root"])
    if model is None:
        model = create_model(**args.unet_model).to(dev).eval()
    rgb_guidance = bool(getattr(args, "rgb_guidance", False))
    paths = None
    if out_dir is not None and getattr(args, "save_singles", False):
        paths = {k: pjoin(out_dir, "single_images", k) for k in ("input", "rgb", "depth_color", "depth_raw")}
        for p in paths.values():
            os.makedirs(p, exist_ok=True)
    loader = datao.ShardedImageLoader(dataset, batch_per_rank, rank=rank, world=world, device=dev,
                                      degamma=False, stop_after=dc.get("stop_after", -1), size=image_size)
    results = []
    for ref_img, names, extras in loader:                       # ref_img [b,3,S,S] in [-1,1], already on the device
        b = ref_img.shape[0]
        opc = dict(args.measurement["operator"]); opc["batch_size"] = b
        operator = get_operator(device=dev, **opc)
        noiser = get_noise(**args.measurement["noise"])
        extra = {} if rgb_guidance else {**args.sample_pattern, **(args.aux_loss or {})}
        cond = get_conditioning_method(args.conditioning["method"], operator, noiser, **args.conditioning["params"], **extra)
        sampler = create_sampler(**args.diffusion)
        y_n = noiser(ref_img)
        if getattr(args, "degamma_input", False):
            y_n = datao.degamma_input(y_n)                      # osmosis_sampling.py:173-175
        pattern = args.sample_pattern["pattern"]
        if pattern not in ("original", "pcgs"):
            raise ValueError(f"Unrecognized sample pattern: {pattern}")
        global_N = 1 if pattern == "original" else args.sample_pattern["global_N"]
        for _ in range(global_N):
            torch.manual_seed(args.manual_seed)                 # the reference reseeds per image: every image sees the same x_T
            x_start = torch.randn(1, 4 if args.unet_model["pretrain_model"] == "osmosis" else 3, image_size, image_size,
                                  device=dev).expand(b, -1, -1, -1).contiguous()
            out = sampler.p_sample_loop(model=model, x_start=x_start, measurement=y_n, measurement_cond_fn=cond.conditioning,
                                        record=False, save_root=None, pretrain_model=args.unet_model["pretrain_model"],
                                        rgb_guidance=rgb_guidance, sample_pattern=args.sample_pattern, noise_mode="shared",
                                        max_steps=max_steps, cuda_graph=cuda_graph)
        if rgb_guidance:
            sample = out
            for k, n in enumerate(names):
                results.append(dict(name=n, loss=float(sampler.last_loss[k])))
                if paths:
                    stem = os.path.splitext(n)[0]
                    _save_png(0.5 * (ref_img[k] + 1), pjoin(paths["input"], stem + ".png"))
                    _save_png(0.5 * (sample[k, :3] + 1), pjoin(paths["rgb"], stem + ".png"))
            continue
        sample, variable_dict, loss, out_xstart = out
        post = utilso.postprocess_samples(operator, out_xstart, ref_img)
        norm_loss = post["norm_loss"].cpu().numpy()
        psnr = None
        if "gt_rgb" in extras:    # simulation sets (ImagesFolder_GT): PSNR of the restored RGB against the ground truth, in [0, 1]
            mse = ((post["sample_rgb_01_clip"] - 0.5 * (extras["gt_rgb"] + 1)) ** 2).mean(dim=(1, 2, 3))
            psnr = (10 * torch.log10(1.0 / mse.clamp_min(1e-12))).cpu().numpy()
        for k, n in enumerate(names):
            r = dict(name=n, loss=float(loss[k]), norm_loss=float(np.round(norm_loss[k], 3)))
            if psnr is not None:
                r["psnr_rgb"] = float(psnr[k])
            for key, v in variable_dict.items():
                r[key] = [round(float(t), 3) for t in v[k].flatten().cpu()]
            results.append(r)
            if paths:
                stem = os.path.splitext(n)[0]
                _save_png(0.5 * (ref_img[k] + 1), pjoin(paths["input"], stem + ".png"))
                _save_png(post["sample_rgb_01_clip"][k], pjoin(paths["rgb"], stem + ".png"))
                _save_png(post["sample_depth_vis_pmm_color"][k], pjoin(paths["depth_color"], stem + ".png"))
                _save_png(post["sample_depth_mm"][k], pjoin(paths["depth_raw"], stem + ".png"))
    return results


def run_prior_sampling(args, device=None, out_dir: str | None = None, model=None, image_size: int | None = None):
    """The body of the reference's `RGBD_prior_sampling.py:main` (:56-122, BASELINE config 1): `number_of_images` unguided RGBD
    samples through `osmosis_utils.diffusion.GaussianDiffusion.inverse`.  Returns a list of dicts {x [1,4,S,S] on the device,
    x_start_rgb, x_start_depth} and writes <out_dir>/{single_images/rgb, single_images/depth, grid_results}/image_<i>.png
    like the reference when `args.save_singles` / `args.save_grids` are set."""
    from .osmosis_utils.diffusion import GaussianDiffusion
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if model is None:
        model = create_model(**args.unet_model).to(dev).eval()
    S = int(image_size if image_size is not None else getattr(args, "image_size", args.unet_model["image_size"]))
    x_start_dim = 4 if args.unet_model["pretrain_model"] == "osmosis" else 3
    rgb_dir = depth_dir = grid_dir = None
    if out_dir is not None:
        rgb_dir, depth_dir, grid_dir = pjoin(out_dir, "single_images", "rgb"), pjoin(out_dir, "single_images", "depth"), pjoin(out_dir, "grid_results")
        for d in (rgb_dir, depth_dir, grid_dir):
            os.makedirs(d, exist_ok=True)
    torch.manual_seed(args.manual_seed)                                   # once, before the loop (:83)
    results = []
    for im_idx in range(args.number_of_images):
        diffusion = GaussianDiffusion(T=args.diffusion["steps"], schedule=args.diffusion["noise_schedule"])
        x, (x_start_rgb, x_start_depth) = diffusion.inverse(
            net=model, shape=(x_start_dim, S, S), image_channels=x_start_dim, steps=args.diffusion["timestep_respacing"], device=dev,
            record_process=getattr(args, "record_process", False), record_every=getattr(args, "record_every", 200),
            save_path=grid_dir, image_idx=im_idx)
        results.append(dict(x=x, x_start_rgb=x_start_rgb, x_start_depth=x_start_depth))
        if out_dir is None:
            continue
        if getattr(args, "save_singles", False) and x_start_rgb is not None:
            _save_png(x_start_rgb, pjoin(rgb_dir, f"image_{im_idx}.png"))
            if x_start_depth is not None:
                _save_png(x_start_depth, pjoin(depth_dir, f"image_{im_idx}.png"))
        if getattr(args, "save_grids", False) and x_start_dim == 4:
            x_rgb = 0.5 * (1 + x[0, 0:3])
            x_d_pmm = utilso.min_max_norm_range_percentile(x[:, 3].contiguous(), percent_low=0.05, percent_high=0.99)
            _save_png(torch.cat([x_rgb, utilso.depth_tensor_to_color_image(x_d_pmm)], dim=2), pjoin(grid_dir, f"image_{im_idx}.png"))
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config_file", default="./configs/osmosis_sample_config.yaml")
    ap.add_argument("-d", "--device", default=0, type=int)
    ap.add_argument("--batch", default=1, type=int, help="images per GPU per sampling run")
    ap.add_argument("--out", default=None)
    ap.add_argument("--prior", action="store_true", help="unguided RGBD prior sampling (RGBD_prior_sampling.py, RGBD_sample_config.yaml)")
    a = ap.parse_args()
    args = utilso.arguments_from_file(os.path.abspath(a.config_file))
    if a.prior:
        torch.cuda.set_device(a.device)
        run_prior_sampling(args, device=f"cuda:{a.device}", out_dir=a.out or pjoin(os.path.abspath(args.save_dir), "rgbd_prior"))
        return
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(a.device)))
    torch.cuda.set_device(local)
    out = a.out or pjoin(os.path.abspath(args.save_dir), args.measurement["operator"]["name"], args.data["name"])
    for r in run_sampling(args, device=f"cuda:{local}", batch_per_rank=a.batch, rank=rank, world=world, out_dir=out):
        print(r)


if __name__ == "__main__":
    main()
