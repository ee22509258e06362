"""Times the GroupNorm kernels alone (CUDA events, median of reps, L2 flushed between reps).  Shapes: B,H,W,C.  Development tool."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402


def main():
    lib = L_.load()
    dev = "cuda"
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1, 256, 256, 256), (8, 256, 256, 256), (8, 128, 128, 512), (1, 64, 64, 512)]
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for (B, H, W, C) in shapes:
        x = torch.randn(B, H, W, C, device=dev); gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
        ss = 0.1 * torch.randn(B, 2 * C, device=dev); stats = torch.zeros(B, 32, 2, device=dev)
        y = torch.empty(B, H, W, C, device=dev); dy = torch.randn(B, H, W, C, device=dev); dx = torch.empty(B, H, W, C, device=dev)
        add = torch.randn(B, H, W, C, device=dev)
        st = L_.stream()
        fwd = lambda: L_.check(lib.osm_dbg_gn_forward(L_.ptr(x), C, L_.ptr(gamma), L_.ptr(beta), L_.ptr(ss), 2 * C, 1, 0, L_.ptr(stats), L_.ptr(y), B, H, W, C, st))
        bwd = lambda: L_.check(lib.osm_dbg_gn_backward(L_.ptr(x), C, L_.ptr(gamma), L_.ptr(beta), L_.ptr(ss), 2 * C, 1, 0, L_.ptr(stats), L_.ptr(dy), L_.ptr(add), C, 1, L_.ptr(dx), C, 0, B, H, W, C, st))
        n = B * H * W * C * 4
        for name, f, byts in (("fwd (stats + apply)", fwd, 3 * n), ("bwd (reduce + apply, +addend)", bwd, 6 * n)):
            ts = []
            for rep in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); torch.cuda.synchronize()
                if rep >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            print(f"{(B, H, W, C)} {name:30s} {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
