"""Development probe of the halo-tile conv kernel (conv_tc_halo_2sm_kernel): one variant per process.
usage: python tools/halo_probe.py TILE_N XFORM [B H W Cin Cout] [--time]"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402


def main():
    tile_n, xform = int(sys.argv[1]), int(sys.argv[2])
    B, H, W, cin, cout = [int(v) for v in sys.argv[3:8]] if len(sys.argv) >= 8 else (2, 32, 32, 64, 256)
    lib = L_.load()
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    bias = torch.randn(cout, generator=g)
    coef = torch.zeros(B, cin, 2)
    coef[..., 0] = 0.5 + torch.rand(B, cin, generator=g)
    coef[..., 1] = torch.randn(B, cin, generator=g) * 0.3
    xa = F.silu(x * coef[..., 0][:, :, None, None] + coef[..., 1][:, :, None, None]) if xform else x
    want = F.conv2d(xa.double(), w.double(), bias.double(), padding=1).float()
    wd = w.contiguous().to(dev)
    wf = torch.zeros(9 * cout * cin, device=dev); wdg = torch.zeros_like(wf)
    L_.check(lib.osm_dbg_pack_conv_weight(L_.ptr(wd), L_.ptr(wf), L_.ptr(wdg), cout, cin, cout, cin, 9, 1, L_.stream()))
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    out = torch.full((B, H, W, cout), float("nan"), device=dev)
    cd = coef.contiguous().to(dev)
    bd = bias.to(dev)
    L_.check(lib.osm_dbg_conv_halo(L_.ptr(xd), cin, L_.ptr(wf), L_.ptr(bd), L_.ptr(cd) if xform else None, 1, None, 0, 0, L_.ptr(out), cout,
                                   B, H, W, cin, cout, tile_n, L_.stream()))
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).cpu()
    err = float((got - want).abs().max() / want.abs().max())
    nan = int(torch.isnan(got).sum())
    # where is it wrong: interior vs border
    d = (got - want).abs().amax(dim=1)[0]
    print(f"tile_n={tile_n} xform={xform} shape=({B},{H},{W},{cin},{cout}): rel err {err:.3e} nan={nan} "
          f"interior {float(d[2:-2, 2:-2].max()):.2e} border {float(d.max()):.2e}  {'OK' if err < 3e-3 and nan == 0 else 'FAIL'}", flush=True)
    if "--time" in sys.argv:
        ts = []
        for i in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.osm_dbg_conv_halo(L_.ptr(xd), cin, L_.ptr(wf), L_.ptr(bd), L_.ptr(cd) if xform else None, 1, None, 0, 0, L_.ptr(out), cout,
                                  B, H, W, cin, cout, tile_n, L_.stream())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        # the unfused tcgen05 path on the same shape, for reference
        ts2 = []
        out2 = torch.empty_like(out)
        for i in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.osm_dbg_conv(0, L_.ptr(xd), cin, L_.ptr(wf), L_.ptr(bd), None, 0, 0, L_.ptr(out2), cout, 0, B, H, W, cin, cout, 9, L_.stream())
            e1.record(); torch.cuda.synchronize()
            ts2.append(e0.elapsed_time(e1))
        t2 = sorted(ts2[2:])[len(ts2[2:]) // 2]
        fl = 2.0 * B * H * W * cin * cout * 9
        print(f"   halo {t*1e3:.1f} us = {fl/t/1e9:.0f} TFLOP/s | current kernel {t2*1e3:.1f} us = {fl/t2/1e9:.0f} TFLOP/s (dbg entry: plans per call)", flush=True)


if __name__ == "__main__":
    main()
