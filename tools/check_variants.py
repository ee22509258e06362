"""Full-size consistency check of the engine's optional fast paths (development tool, GPU): builds the shipped config's UNet
several times with the same synthetic weights and compares UNet output and input-VJP between
  * fused tcgen05 attention (OSM_ATTN_FLASH=1) vs the fp32-accurate batched-GEMM attention (=0),
  * GroupNorm statistics reduced in the conv epilogues (OSM_GN_FUSE=1) vs the stand-alone kernels (=0).
Both switches are read when a model is created.   python tools/check_variants.py [--batch 1]
"""
import argparse
import contextlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file  # noqa: E402
from osmosis_diffusion_code_b200.guided_diffusion.unet import create_model  # noqa: E402
from osmosis_diffusion_code_b200.synthetic import synth_state_dict  # noqa: E402


def run(env, um, B, size, x, t, g):
    os.environ.update(env)
    with contextlib.redirect_stdout(sys.stderr):
        model = create_model(**um)
    sd = synth_state_dict(model.param_specs(), um["num_channels"], seed=7, delta=0.05)
    model.load_state_dict(sd); del sd
    model.to("cuda")
    out = model._forward_raw(x, t).clone()
    gx = model._vjp_raw(g).clone()
    torch.cuda.synchronize()
    n = model.launch_counts()
    del model
    torch.cuda.empty_cache()
    return out, gx, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--size", type=int, default=256)
    args = ap.parse_args()
    a = arguments_from_file("configs/osmosis_sample_config.yaml")
    um = dict(a.unet_model); um["model_path"] = ""
    B, S = args.batch, args.size
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, 4, S, S, generator=gen).cuda()
    t = torch.full((B,), 500.0).cuda()
    g = (torch.randn(B, 8, S, S, generator=gen) * 1e-3).cuda()
    base = run({"OSM_ATTN_FLASH": "0", "OSM_GN_FUSE": "0"}, um, B, S, x, t, g)
    ok = True
    for name, env in (("flash attention", {"OSM_ATTN_FLASH": "1", "OSM_GN_FUSE": "0"}),
                      ("fused GroupNorm statistics (fwd)", {"OSM_ATTN_FLASH": "0", "OSM_GN_FUSE": "1"}),
                      ("fused GroupNorm statistics (fwd+bwd)", {"OSM_ATTN_FLASH": "0", "OSM_GN_FUSE": "2"}),
                      ("flash + fused fwd stats (default)", {"OSM_ATTN_FLASH": "1", "OSM_GN_FUSE": "1"})):
        o, gx, n = run(env, um, B, S, x, t, g)
        eo = float((o - base[0]).abs().max() / base[0].abs().max())
        eg = float((gx - base[1]).abs().max() / base[1].abs().max())
        fin = bool(torch.isfinite(o).all() and torch.isfinite(gx).all())
        good = fin and eo < 5e-3 and eg < 2e-2
        ok &= good
        print(f"{name:38s}: out diff {eo:.2e}  input-VJP diff {eg:.2e}  finite={fin}  launches fwd/vjp {n} (base {base[2]})  {'OK' if good else 'FAIL'}")
    print("ALL OK" if ok else "SOME FAILED")


if __name__ == "__main__":
    main()
