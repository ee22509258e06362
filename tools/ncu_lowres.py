"""Stand-alone launches of the batch-1 low-resolution kernels for an `ncu --set full` capture: the cluster split-K conv (fp16 operands from
memory, partials reduced through the L2 scratch) and the one-launch fp32 attention of the 8x8 level.  Two launches each: the second is the
one to read.

    ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|attn_small" -o gpurun_out/ncu_lowres python tools/ncu_lowres.py
"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402

SHAPES = [(1, 8, 8, 1024, 1024, 9), (1, 16, 16, 1024, 1024, 9), (1, 32, 32, 512, 512, 9), (1, 8, 8, 1024, 1024, 1)]


def main():
    lib = L_.load()
    dev = "cuda"
    for (B, H, W, cin, cout, taps) in SHAPES:
        k = 3 if taps == 9 else 1
        g = torch.Generator().manual_seed(1)
        w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * taps)).to(dev)
        wf = torch.zeros(taps * cout * cin // 2, device=dev); wd = torch.zeros_like(wf)
        L_.check(lib.osm_dbg_pack_conv_weight_f16(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, L_.stream()))
        x = torch.randn(B, H, W, cin, device=dev).half(); bias = torch.randn(cout, device=dev); out = torch.empty(B, H, W, cout, device=dev)
        for _ in range(2):
            L_.check(lib.osm_dbg_conv_f16(C.c_void_p(x.data_ptr()), cin, L_.ptr(wf), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B, H, W, cin, cout,
                                          taps, L_.stream()))
        torch.cuda.synchronize()
        print("ran conv", (B, H, W, cin, cout, taps), float(out.abs().mean()))
    B, L, Cc, heads = 1, 64, 1024, 16
    qkv = torch.randn(B, L, 3 * Cc, device=dev); go = torch.randn(B, L, Cc, device=dev)
    out = torch.zeros(B, L, Cc, device=dev); gq = torch.zeros(B, L, 3 * Cc, device=dev)
    P = torch.zeros(B * heads * L * L, device=dev); D = torch.zeros_like(P)
    for _ in range(2):
        L_.check(lib.osm_dbg_attention(L_.ptr(qkv), L_.ptr(out), L_.ptr(P), B, L, Cc, heads, L_.stream()))
        L_.check(lib.osm_dbg_attention_bwd(L_.ptr(qkv), L_.ptr(go), L_.ptr(gq), L_.ptr(P), L_.ptr(D), B, L, Cc, heads, L_.stream()))
    torch.cuda.synchronize()
    print("ran attention", (B, L, Cc, heads), float(out.abs().mean()), float(gq.abs().mean()))


if __name__ == "__main__":
    main()
