"""Per-block time of the attention paths inside a CUDA graph (20 dependent forward / backward blocks replayed): the fused tcgen05
path (transpose + flash kernels) against the fp32 path (one-launch attn_small_* kernels at L = 64).  Development tool.

    python tools/time_attn.py [B,L,C,heads ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402


def graph_us(f, reps=20):
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            f()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


def run(lib, B, L, C, heads):
    dev = "cuda"
    qkv = torch.randn(B, L, 3 * C, device=dev)
    go = torch.randn(B, L, C, device=dev)
    out = torch.zeros(B, L, C, device=dev); gq = torch.zeros(B, L, 3 * C, device=dev)
    qkvT = torch.zeros(B, 3 * C, L, device=dev); goT = torch.zeros(B, C, L, device=dev)
    lse = torch.zeros(B, heads, L, device=dev); Dv = torch.zeros(B, heads, L, device=dev)
    P = torch.zeros(B * heads * L * L, device=dev); D = torch.zeros_like(P)
    st = L_.stream
    t = {}
    t["flash fwd"] = graph_us(lambda: L_.check(lib.osm_dbg_attention_flash(L_.ptr(qkv), L_.ptr(qkvT), L_.ptr(out), L_.ptr(lse), B, L, C, heads, st())))
    t["flash bwd"] = graph_us(lambda: L_.check(lib.osm_dbg_attention_flash_bwd(L_.ptr(qkv), L_.ptr(qkvT), L_.ptr(out), L_.ptr(lse), L_.ptr(Dv), L_.ptr(go),
                                                                               L_.ptr(goT), L_.ptr(gq), B, L, C, heads, st())))
    t["fp32 fwd"] = graph_us(lambda: L_.check(lib.osm_dbg_attention(L_.ptr(qkv), L_.ptr(out), L_.ptr(P), B, L, C, heads, st())))
    t["fp32 bwd"] = graph_us(lambda: L_.check(lib.osm_dbg_attention_bwd(L_.ptr(qkv), L_.ptr(go), L_.ptr(gq), L_.ptr(P), L_.ptr(D), B, L, C, heads, st())))
    print(f"{(B, L, C, heads)}: " + "  ".join(f"{k} {v:.2f} us" for k, v in t.items()), flush=True)


if __name__ == "__main__":
    lib = L_.load()
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1, 64, 1024, 16), (8, 64, 1024, 16), (32, 64, 1024, 16)]
    for sh in shapes:
        run(lib, *sh)
