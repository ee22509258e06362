"""Per-launch time of the tcgen05 conv inside a CUDA graph (50 dependent launches replayed, warm L2): the launch floor and
the small-layer times the step actually sees.  Development tool.   python tools/conv_floor.py [B,H,W,Cin,Cout,taps ...]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402

DEFAULT = [(1, 8, 8, 64, 128, 1), (1, 8, 8, 1024, 1024, 9), (1, 8, 8, 1024, 1024, 1), (1, 16, 16, 1024, 1024, 9), (1, 16, 16, 1024, 1024, 1),
           (1, 32, 32, 512, 512, 9), (1, 64, 64, 512, 512, 9), (1, 128, 128, 256, 256, 9), (1, 256, 256, 256, 256, 9)]


def run(lib, B, H, W, cin, cout, taps, reps=50):
    dev = "cuda"
    k = 3 if taps == 9 else 1
    w = (torch.randn(cout, cin, k, k) / math.sqrt(cin * taps)).to(dev)
    wf = torch.zeros(taps * cout * cin, device=dev); wd = torch.zeros_like(wf)
    L_.check(lib.osm_dbg_pack_conv_weight(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, 1, L_.stream()))
    x = torch.randn(B, H, W, cin, device=dev); bias = torch.randn(cout, device=dev); out = torch.empty(B, H, W, cout, device=dev)
    f = lambda: L_.check(lib.osm_dbg_conv(0, L_.ptr(x), cin, L_.ptr(wf), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B, H, W, cin, cout,
                                          taps, L_.stream()))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            f()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    print(f"{(B, H, W, cin, cout, taps)}: {best:.1f} us per launch inside a graph")


if __name__ == "__main__":
    lib = L_.load()
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or DEFAULT
    for sh in shapes:
        run(lib, *sh)
