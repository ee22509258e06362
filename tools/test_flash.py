"""Development check of the fused tcgen05 attention kernels against torch (fp64 reference on the GPU) + timing against the
round-1 bgemm path.  Usage: python tools/test_flash.py [B,L,heads ...]"""
import math
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402

DEFAULT = ["1,64,16", "2,64,16", "1,256,16", "1,1024,8", "2,1024,8", "8,1024,8"]


def ref(qkv, go, heads):
    B, L, C3 = qkv.shape
    C = C3 // 3
    q = qkv.double().detach().requires_grad_()
    x = q.view(B, L, heads, 3, 64)
    qq, kk, vv = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]            # [B, L, h, 64]
    s = torch.einsum("bthc,bshc->bhts", qq, kk) / math.sqrt(64.0)
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhts,bshc->bthc", p, vv).reshape(B, L, C)
    (g,) = torch.autograd.grad(o, q, go.double())
    return o.float(), g.float()


def timeit(fn, reps=10):
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) * 1e3


def main():
    lib = L_.load()
    dev = "cuda"
    ok = True
    for sh in (sys.argv[1:] or DEFAULT):
        B, L, heads = [int(v) for v in sh.split(",")]
        C = heads * 64
        g = torch.Generator().manual_seed(L + B)
        # peaked-ish logits: scale q,k so that scores have std ~ 2
        qkv = (torch.randn(B, L, 3 * C, generator=g) * 1.2).to(dev)
        go = torch.randn(B, L, C, generator=g).to(dev)
        o_ref, g_ref = ref(qkv, go, heads)
        qkvT = torch.zeros(B, 3 * C, L, device=dev)
        out = torch.zeros(B, L, C, device=dev)
        lse = torch.zeros(B, heads, L, device=dev)
        Dv = torch.zeros(B, heads, L, device=dev)
        goT = torch.zeros(B, C, L, device=dev)
        gq = torch.zeros(B, L, 3 * C, device=dev)
        st = L_.stream()
        fwd = lambda: L_.check(lib.osm_dbg_attention_flash(L_.ptr(qkv), L_.ptr(qkvT), L_.ptr(out), L_.ptr(lse), B, L, C, heads, st))
        bwd = lambda: L_.check(lib.osm_dbg_attention_flash_bwd(L_.ptr(qkv), L_.ptr(qkvT), L_.ptr(out), L_.ptr(lse), L_.ptr(Dv), L_.ptr(go),
                                                               L_.ptr(goT), L_.ptr(gq), B, L, C, heads, st))
        fwd(); torch.cuda.synchronize()
        bwd(); torch.cuda.synchronize()
        eo = float((out - o_ref).abs().max() / o_ref.abs().max())
        x = gq.view(B, L, heads, 3, 64); xr = g_ref.view(B, L, heads, 3, 64)
        errs = [float((x[:, :, :, i] - xr[:, :, :, i]).abs().max() / xr[:, :, :, i].abs().max()) for i in range(3)]
        # lse check
        s = torch.einsum("bthc,bshc->bhts", qkv.view(B, L, heads, 3, 64)[:, :, :, 0].double(), qkv.view(B, L, heads, 3, 64)[:, :, :, 1].double()) / 8.0
        lse_ref = (torch.logsumexp(s, dim=-1) / math.log(2.0)).float()
        el = float((lse - lse_ref).abs().max())
        # old path
        P = torch.empty(B * heads * L * L, device=dev); D = torch.empty_like(P)
        out2 = torch.zeros_like(out); gq2 = torch.zeros_like(gq)
        fwd_old = lambda: L_.check(lib.osm_dbg_attention(L_.ptr(qkv), L_.ptr(out2), L_.ptr(P), B, L, C, heads, st))
        bwd_old = lambda: L_.check(lib.osm_dbg_attention_bwd(L_.ptr(qkv), L_.ptr(go), L_.ptr(gq2), L_.ptr(P), L_.ptr(D), B, L, C, heads, st))
        fwd_old(); bwd_old(); torch.cuda.synchronize()
        eo_old = float((out2 - o_ref).abs().max() / o_ref.abs().max())
        eg_old = float((gq2 - g_ref).abs().max() / g_ref.abs().max())
        tf, tb, tfo, tbo = timeit(fwd), timeit(bwd), timeit(fwd_old), timeit(bwd_old)
        good = eo < 3e-3 and max(errs) < 6e-3 and el < 1e-2
        ok &= good
        print(f"B={B} L={L} heads={heads}: out err {eo:.2e} (old {eo_old:.2e})  dq/dk/dv err {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e} "
              f"(old {eg_old:.2e})  lse err {el:.2e}  | fwd {tf:.1f} us (old {tfo:.1f})  bwd {tb:.1f} us (old {tbo:.1f})  "
              f"{'OK' if good else 'FAIL'}", flush=True)
        if not good:
            # where is it wrong?  per-head / per-row-block error map of the forward output
            d = (out - o_ref).abs().view(B, L // 64, 64, heads, 64).amax(dim=(2, 4))
            print("  fwd err by [b, row block, head] (max):", d[0, :4, :4].tolist())
            nan = [int(torch.isnan(t).sum()) for t in (out, gq, lse, Dv)]
            print("  NaN counts out/gq/lse/Dv:", nan)
    print("ALL OK" if ok else "SOME FAILED")


if __name__ == "__main__":
    main()
