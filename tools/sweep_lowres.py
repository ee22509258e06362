"""Development sweep: (BN, split) of the tile-per-CTA / persistent tcgen05 kernels on fp16 operands for the low-resolution shapes of a
batch-1 step (warm L2, CUDA events, median).  Prints the policy's own choice and the best forced one."""
import math
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402

SHAPES = ["1,8,8,1024,1024,9", "1,8,8,2048,1024,9", "1,16,16,1024,1024,9", "1,16,16,2048,1024,9", "1,32,32,512,512,9", "1,32,32,1024,512,9",
          "1,32,32,1024,1024,9", "1,16,16,1024,3072,1", "1,8,8,1024,3072,1", "1,32,32,512,1536,1", "1,16,16,1024,1024,1", "1,32,32,512,512,1"]


def main():
    lib = L_.load()
    dev = "cuda"
    for sh in sys.argv[1:] or SHAPES:
        B, H, W, cin, cout, taps = [int(t) for t in sh.split(",")]
        k = 3 if taps == 9 else 1
        g = torch.Generator().manual_seed(1)
        w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * taps)).to(dev)
        wf16 = torch.zeros(taps * cout * cin, dtype=torch.float16, device=dev); wd16 = torch.zeros_like(wf16)
        L_.check(lib.osm_dbg_pack_conv_weight_f16(L_.ptr(w), L_.ptr(wf16), L_.ptr(wd16), cout, cin, cout, cin, taps, L_.stream()))
        xh = torch.randn(B, H, W, cin, device=dev).half()
        bias = torch.randn(cout, device=dev)
        out = torch.empty(B, H, W, cout, device=dev)

        def t_of():
            ts = []
            for rep in range(11):
                torch.cuda._sleep(200000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.osm_dbg_conv_f16(L_.ptr(xh), cin, L_.ptr(wf16), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B, H, W, cin, cout, taps, L_.stream())
                e1.record(); torch.cuda.synchronize()
                if rc != 0:
                    return None
                if rep >= 3:
                    ts.append(e0.elapsed_time(e1))
            return statistics.median(ts) * 1e3
        os.environ.pop("OSM_CONV_FORCE", None)
        base = t_of()
        res = []
        for bn in (64, 128, 256):
            if cout % bn:
                continue
            for sp in (1, 2, 4, 8, 16):
                os.environ["OSM_CONV_FORCE"] = f"{bn},{sp}"
                t = t_of()
                if t is not None:
                    res.append((t, bn, sp))
        os.environ.pop("OSM_CONV_FORCE", None)
        res.sort()
        print(f"{sh:24s} policy {base:6.1f} us | best " + "  ".join(f"({bn},{sp}) {t:5.1f}" for t, bn, sp in res[:4]), flush=True)


if __name__ == "__main__":
    main()
