"""Stand-alone launches of the tcgen05 conv kernel for `ncu --set full` captures (one launch per listed shape, after a warm-up
launch of each).  Shapes: B,H,W,Cin,Cout,taps.

    ncu --set full --clock-control none --import-source on -k regex:conv_tc -o gpurun_out/conv python tools/ncu_conv.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402

SHAPES = [(1, 256, 256, 256, 256, 9), (1, 8, 8, 1024, 1024, 9), (8, 256, 256, 256, 256, 9)]


def main():
    lib = L_.load()
    dev = "cuda"
    shapes = SHAPES if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for (B, H, W, cin, cout, taps) in shapes:
        k = 3 if taps == 9 else 1
        g = torch.Generator().manual_seed(1)
        w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * taps)).to(dev)
        wf = torch.zeros(taps * cout * cin, device=dev); wd = torch.zeros_like(wf)
        L_.check(lib.osm_dbg_pack_conv_weight(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, 1, L_.stream()))
        x = torch.randn(B, H, W, cin, device=dev)
        bias = torch.randn(cout, device=dev)
        out = torch.empty(B, H, W, cout, device=dev)
        for _ in range(2):  # first = warm-up (function attributes, descriptors), second = the one to read in the report
            L_.check(lib.osm_dbg_conv(0, L_.ptr(x), cin, L_.ptr(wf), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B, H, W, cin, cout,
                                      taps, L_.stream()))
        torch.cuda.synchronize()
        print("ran", (B, H, W, cin, cout, taps), float(out.abs().mean()))


if __name__ == "__main__":
    main()
