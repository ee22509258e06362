"""Development probe: low-resolution convs with their weights cold (L2 flushed) vs already resident in L2.
Answers whether prefetching the next layer's weights into L2 would pay at batch 1."""
import math
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402

SHAPES = ["1,8,8,1024,1024,9", "1,8,8,2048,1024,9", "1,16,16,1024,1024,9", "1,32,32,512,512,9", "1,32,32,1024,1024,9", "1,16,16,1024,3072,1",
          "1,64,64,512,512,9"]


def main():
    lib = L_.load()
    dev = "cuda"
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for sh in sys.argv[1:] or SHAPES:
        B, H, W, cin, cout, taps = [int(t) for t in sh.split(",")]
        k = 3 if taps == 9 else 1
        g = torch.Generator().manual_seed(1)
        w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * taps)).to(dev)
        wf = torch.zeros(taps * cout * cin, device=dev); wd = torch.zeros_like(wf)
        L_.check(lib.osm_dbg_pack_conv_weight(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, 1, L_.stream()))
        wf16 = torch.zeros(taps * cout * cin, dtype=torch.float16, device=dev); wd16 = torch.zeros_like(wf16)
        L_.check(lib.osm_dbg_pack_conv_weight_f16(L_.ptr(w), L_.ptr(wf16), L_.ptr(wd16), cout, cin, cout, cin, taps, L_.stream()))
        x = torch.randn(B, H, W, cin, device=dev); xh = x.half()
        bias = torch.randn(cout, device=dev)
        out = torch.empty(B, H, W, cout, device=dev)
        row = []
        for name, call in (("tf32", lambda: lib.osm_dbg_conv(0, L_.ptr(x), cin, L_.ptr(wf), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B, H, W, cin, cout, taps, L_.stream())),
                           ("fp16", lambda: lib.osm_dbg_conv_f16(L_.ptr(xh), cin, L_.ptr(wf16), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B, H, W, cin, cout, taps, L_.stream()))):
            for mode in ("cold", "warm"):
                ts = []
                for rep in range(14):
                    if mode == "cold":
                        flush.zero_()
                    else:
                        torch.cuda._sleep(300000)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); L_.check(call()); e1.record(); torch.cuda.synchronize()
                    if rep >= 3:
                        ts.append(e0.elapsed_time(e1))
                row.append(f"{name} {mode} {statistics.median(ts)*1e3:6.1f} us")
        print(f"{sh:24s} " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
