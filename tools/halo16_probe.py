"""Development probe of the fp16-operand halo conv kernel (conv_tc_halo16_2sm_kernel).
usage: python tools/halo16_probe.py MODE [B H W Cin Cout] [--time]     MODE: 0 convert only, 1 affine, 2 affine + SiLU"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402


def main():
    mode = int(sys.argv[1])
    B, H, W, cin, cout = [int(v) for v in sys.argv[2:7]] if len(sys.argv) >= 7 else (2, 32, 32, 64, 256)
    lib = L_.load()
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    bias = torch.randn(cout, generator=g)
    coef = torch.zeros(B, cin, 2)
    coef[..., 0] = 0.5 + torch.rand(B, cin, generator=g)
    coef[..., 1] = torch.randn(B, cin, generator=g) * 0.3
    xa = x
    if mode:
        xa = x * coef[..., 0][:, :, None, None] + coef[..., 1][:, :, None, None]
        if mode == 2:
            xa = F.silu(xa)
    big = B * H * W * cin * cout > 2e10
    wd = w.contiguous().to(dev)
    want = (F.conv2d(xa.to(dev), wd, bias.to(dev), padding=1).cpu() if big else F.conv2d(xa.double(), w.double(), bias.double(), padding=1).float())
    wf16 = torch.zeros(9 * cout * cin, dtype=torch.float16, device=dev); wd16 = torch.zeros_like(wf16)
    L_.check(lib.osm_dbg_pack_conv_weight_f16(L_.ptr(wd), L_.ptr(wf16), L_.ptr(wd16), cout, cin, cout, cin, 9, L_.stream()))
    wf = torch.zeros(9 * cout * cin, device=dev); wdg = torch.zeros_like(wf)
    L_.check(lib.osm_dbg_pack_conv_weight(L_.ptr(wd), L_.ptr(wf), L_.ptr(wdg), cout, cin, cout, cin, 9, 1, L_.stream()))
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    out = torch.full((B, H, W, cout), float("nan"), device=dev)
    cd = coef.contiguous().to(dev)
    bd = bias.to(dev)

    def call():
        return lib.osm_dbg_conv_halo16(L_.ptr(xd), cin, L_.ptr(wf16), L_.ptr(bd), L_.ptr(cd) if mode else None, 1 if mode == 2 else 0, None, 0, 0,
                                       L_.ptr(out), cout, 0, B, H, W, cin, cout, L_.stream())
    L_.check(call())
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).cpu()
    err = float((got - want).abs().max() / want.abs().max())
    nan = int(torch.isnan(got).sum())
    d = (got - want).abs().amax(dim=1)[0]
    print(f"halo16 mode={mode} shape=({B},{H},{W},{cin},{cout}): rel err {err:.3e} nan={nan} "
          f"interior {float(d[2:-2, 2:-2].max()):.2e} border {float(d.max()):.2e}  {'OK' if err < 3e-3 and nan == 0 else 'FAIL'}", flush=True)
    if "--time" in sys.argv:
        def med(fn):
            ts = []
            for i in range(12):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sorted(ts[2:])[len(ts[2:]) // 2]
        t = med(call)
        out2 = torch.empty_like(out)
        t2 = med(lambda: lib.osm_dbg_conv_halo(L_.ptr(xd), cin, L_.ptr(wf), L_.ptr(bd), L_.ptr(cd) if mode else None, 1 if mode == 2 else 0, None, 0, 0,
                                               L_.ptr(out2), cout, B, H, W, cin, cout, 256, L_.stream()))
        fl = 2.0 * B * H * W * cin * cout * 9
        print(f"   halo16 {t*1e3:.1f} us = {fl/t/1e9:.0f} TFLOP/s | tf32 halo kernel {t2*1e3:.1f} us = {fl/t2/1e9:.0f} TFLOP/s (dbg entries plan per call)", flush=True)


if __name__ == "__main__":
    main()
