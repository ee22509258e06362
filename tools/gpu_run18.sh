#!/bin/bash
# fixed cost of a conv launch: tiny-K shapes, no cluster / cluster 2 / 4 / 8, persistent vs one tile per CTA
T="timeout 120 python tools/time_conv.py"
echo "== persistent, K=2 blocks";            OSM_CONV_FORCE=128,1 $T 1,8,8,64,128,1 1,16,16,64,1024,1 2>&1 | grep us
echo "== one tile per CTA, no cluster, K=2";  OSM_CONV_PERSIST=0 OSM_CONV_FORCE=128,1 $T 1,8,8,64,128,1 1,16,16,64,1024,1 2>&1 | grep us
echo "== cluster 2, K=2 blocks";             OSM_CONV_FORCE=128,2 $T 1,8,8,64,128,1 1,16,16,64,1024,1 2>&1 | grep us
echo "== cluster 4, K=4 blocks";             OSM_CONV_FORCE=128,4 $T 1,8,8,128,128,1 1,16,16,128,1024,1 2>&1 | grep us
echo "== cluster 8, K=8 blocks";             OSM_CONV_FORCE=128,8 $T 1,8,8,256,128,1 1,16,16,256,1024,1 2>&1 | grep us
echo "== cluster 8, K=288 blocks (8x8 1024->1024 3x3), BN=128 / 256"; OSM_CONV_FORCE=128,8 $T 1,8,8,1024,1024,9 2>&1 | grep us; OSM_CONV_FORCE=256,8 $T 1,8,8,1024,1024,9 2>&1 | grep us
echo "== cluster 8, K=288, Cout=128 only (1 N tile: 8 CTAs)"; OSM_CONV_FORCE=128,8 $T 1,8,8,1024,128,9 2>&1 | grep us
echo "== no-flush back-to-back (graph-like): 20 launches"; 
python - <<'PY'
import math, os, sys, torch
sys.path.insert(0, os.getcwd())
from osmosis_diffusion_code_b200 import lib as L_
lib = L_.load(); dev = "cuda"
def run(B,H,W,cin,cout,taps,reps=50):
    k = 3 if taps == 9 else 1
    w = (torch.randn(cout, cin, k, k) / math.sqrt(cin*taps)).to(dev)
    wf = torch.zeros(taps*cout*cin, device=dev); wd = torch.zeros_like(wf)
    L_.check(lib.osm_dbg_pack_conv_weight(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, 1, L_.stream()))
    x = torch.randn(B,H,W,cin, device=dev); bias = torch.randn(cout, device=dev); out = torch.empty(B,H,W,cout, device=dev)
    f = lambda: L_.check(lib.osm_dbg_conv(0, L_.ptr(x), cin, L_.ptr(wf), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B,H,W,cin,cout,taps, L_.stream()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{(B,H,W,cin,cout,taps)}: {e0.elapsed_time(e1)/reps*1e3:.1f} us per launch inside a graph (warm L2, dependent chain)")
for sh in [(1,8,8,64,128,1),(1,8,8,1024,1024,9),(1,8,8,1024,1024,1),(1,16,16,1024,1024,9),(1,16,16,1024,1024,1),(1,32,32,512,512,9),(1,64,64,512,512,9),(1,128,128,256,256,9),(1,256,256,256,256,9)]:
    run(*sh)
PY
