"""Times the tcgen05 conv kernel alone on a list of shapes (CUDA events, median of reps; a 256 MB buffer is written between
reps to flush L2).  Shapes: B,H,W,Cin,Cout,taps[,res]   Development tool."""
import math
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402

DEFAULT = ["1,256,256,256,256,9", "1,256,256,256,256,9,1", "8,256,256,256,256,9", "8,256,256,256,256,9,1", "1,256,256,512,256,1",
           "8,256,256,512,256,1", "1,128,128,256,256,9", "8,128,128,512,512,9", "1,64,64,512,512,9", "1,8,8,1024,1024,9",
           "1,16,16,1024,1024,9", "1,32,32,512,512,9", "8,8,8,1024,1024,9"]


def main():
    lib = L_.load()
    dev = "cuda"
    shapes = sys.argv[1:] or DEFAULT
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for sh in shapes:
        v = [int(t) for t in sh.split(",")]
        B, H, W, cin, cout, taps = v[:6]
        res_mode = v[6] if len(v) > 6 else 0
        stat_mode = v[7] if len(v) > 7 else 0   # 1 / 2: GroupNorm statistics reduced in the epilogue (+ finalize kernel)
        k = 3 if taps == 9 else 1
        g = torch.Generator().manual_seed(1)
        w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * taps)).to(dev)
        wf = torch.zeros(taps * cout * cin, device=dev); wd = torch.zeros_like(wf)
        L_.check(lib.osm_dbg_pack_conv_weight(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, 1, L_.stream()))
        x = torch.randn(B, H, W, cin, device=dev)
        bias = torch.randn(cout, device=dev)
        res = torch.randn(B, H, W, cout, device=dev) if res_mode else None
        out = torch.empty(B, H, W, cout, device=dev)
        ts = []
        if stat_mode:
            import ctypes
            gx = torch.randn(B, H, W, cout, device=dev); gam = torch.ones(cout, device=dev); bet = torch.zeros(cout, device=dev)
            fst = torch.zeros(B, 32, 2, device=dev); fst[..., 1] = 1.0
            part = torch.empty(B * (H * W // 128 + 64) * 256, device=dev); coef = torch.empty(B * cout * 4, device=dev)
            so = torch.zeros(B, 32, 2, device=dev); fused = ctypes.c_int(0)
        for rep in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if stat_mode:
                L_.check(lib.osm_dbg_conv_stats(L_.ptr(x), cin, L_.ptr(wf), L_.ptr(bias), L_.ptr(out), cout, B, H, W, cin, cout, taps,
                                                stat_mode, L_.ptr(gx), cout, L_.ptr(gam), L_.ptr(bet), None, 0, 1, L_.ptr(fst),
                                                L_.ptr(part), L_.ptr(coef), L_.ptr(so), ctypes.addressof(fused), L_.stream()))
            else:
                L_.check(lib.osm_dbg_conv(0, L_.ptr(x), cin, L_.ptr(wf), L_.ptr(bias), L_.ptr(res), cout, res_mode, L_.ptr(out), cout, 0,
                                          B, H, W, cin, cout, taps, L_.stream()))
            e1.record(); torch.cuda.synchronize()
            if rep >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        fl = 2.0 * B * H * W * cin * cout * taps
        print(f"{sh:28s} {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TF/s")


if __name__ == "__main__":
    main()
