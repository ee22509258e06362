"""Development probe: the fp16 halo conv kernel with the fused GroupNorm statistics epilogue (mode 1 forward / 2 backward), timed.
usage: python tools/halo16_stats_probe.py MODE [B H W Cin Cout]"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200 import lib as L_  # noqa: E402


def main():
    mode = int(sys.argv[1])
    B, H, W, cin, cout = [int(v) for v in sys.argv[2:7]] if len(sys.argv) >= 7 else (2, 256, 256, 256, 256)
    lib = L_.load()
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, W, cin, generator=g).to(dev)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(dev)
    wf = torch.zeros(9 * cout * cin, dtype=torch.float16, device=dev); wd = torch.zeros_like(wf)
    L_.check(lib.osm_dbg_pack_conv_weight_f16(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, 9, L_.stream()))
    out = torch.empty(B, H, W, cout, device=dev)
    part = torch.zeros(B * (H // 16) * (W // 8) * 4 * 64, device=dev)
    coef = torch.empty(B * cout * 4, device=dev)
    stats = torch.zeros(B, 32, 2, device=dev)
    gx = torch.randn(B, H, W, cout, generator=g).to(dev)
    gamma = torch.ones(cout, device=dev); beta = torch.zeros(cout, device=dev)
    fstats = torch.zeros(B, 32, 2, device=dev); fstats[..., 1] = 1.0
    fused = C.c_int(0)

    def call():
        L_.check(lib.osm_dbg_conv_stats_f16(L_.ptr(x), cin, L_.ptr(wf), None, L_.ptr(out), cout, B, H, W, cin, cout, 9, mode, L_.ptr(gx), cout,
                                            L_.ptr(gamma), L_.ptr(beta), None, 0, 1, L_.ptr(fstats), L_.ptr(part), L_.ptr(coef), L_.ptr(stats),
                                            C.addressof(fused), L_.stream()))
    call(); torch.cuda.synchronize()
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[2:])[len(ts[2:]) // 2]
    fl = 2.0 * B * H * W * cin * cout * 9
    print(f"halo16 stats mode {mode} ({B},{H},{W},{cin},{cout}): {t*1e3:.1f} us (conv + coef + finalize) = {fl/t/1e9:.0f} TFLOP/s fused={fused.value}", flush=True)


if __name__ == "__main__":
    main()
