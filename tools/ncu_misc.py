"""Stand-alone launches of the GroupNorm kernels and the fused attention kernels at the shipped config's shapes, for
`ncu --set full` captures (each kernel is launched twice: warm-up, then the launch to read in the report).

    ncu --set full --clock-control none --import-source on -k regex:"gn_|flash_" -o gpurun_out/misc python tools/ncu_misc.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402


def gn(lib, B, H, W, C):
    dev = "cuda"
    x = torch.randn(B, H, W, C, device=dev)
    gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    ss = 0.1 * torch.randn(B, 2 * C, device=dev)
    stats = torch.zeros(B, 32, 2, device=dev)
    y = torch.empty(B, H, W, C, device=dev)
    dy = torch.randn(B, H, W, C, device=dev)
    dx = torch.empty(B, H, W, C, device=dev)
    st = L_.stream()
    for _ in range(2):
        L_.check(lib.osm_dbg_gn_forward(L_.ptr(x), C, L_.ptr(gamma), L_.ptr(beta), L_.ptr(ss), 2 * C, 1, 0, L_.ptr(stats), L_.ptr(y),
                                        B, H, W, C, st))
        L_.check(lib.osm_dbg_gn_backward(L_.ptr(x), C, L_.ptr(gamma), L_.ptr(beta), L_.ptr(ss), 2 * C, 1, 0, L_.ptr(stats), L_.ptr(dy),
                                         None, C, 0, L_.ptr(dx), C, 0, B, H, W, C, st))
    torch.cuda.synchronize()
    print("gn", (B, H, W, C), float(y.abs().mean()), float(dx.abs().mean()))


def flash(lib, B, L, heads):
    dev = "cuda"
    C = heads * 64
    qkv = torch.randn(B, L, 3 * C, device=dev)
    go = torch.randn(B, L, C, device=dev)
    qkvT = torch.zeros(B, 3 * C, L, device=dev); out = torch.zeros(B, L, C, device=dev)
    lse = torch.zeros(B, heads, L, device=dev); Dv = torch.zeros(B, heads, L, device=dev)
    goT = torch.zeros(B, C, L, device=dev); gq = torch.zeros(B, L, 3 * C, device=dev)
    st = L_.stream()
    for _ in range(2):
        L_.check(lib.osm_dbg_attention_flash(L_.ptr(qkv), L_.ptr(qkvT), L_.ptr(out), L_.ptr(lse), B, L, C, heads, st))
        L_.check(lib.osm_dbg_attention_flash_bwd(L_.ptr(qkv), L_.ptr(qkvT), L_.ptr(out), L_.ptr(lse), L_.ptr(Dv), L_.ptr(go), L_.ptr(goT),
                                                 L_.ptr(gq), B, L, C, heads, st))
    torch.cuda.synchronize()
    print("flash", (B, L, heads), float(out.abs().mean()), float(gq.abs().mean()))


def main():
    lib = L_.load()
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if which in ("all", "gn"):
        gn(lib, B, 256, 256, 256)
        gn(lib, B, 32, 32, 512)
    if which in ("all", "flash"):
        flash(lib, B, 1024, 8)
        flash(lib, B, 256, 16)


if __name__ == "__main__":
    main()
