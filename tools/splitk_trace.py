"""Phase timeline of the cluster split-K conv kernel (conv_tc_kernel) inside a CUDA graph.  Development tool.

    python tools/splitk_trace.py --build            # here (no GPU): compile conv_tc.cu with -DOSM_TRACE, link tools/_trace/*.so
    python tools/splitk_trace.py [B,H,W,Cin,Cout,taps ...]     # on the GPU box
    python tools/splitk_trace.py --sweep [shapes ...]          # forced (BN, split) sweep, in-graph time with COLD weights

For every shape: 30 dependent launches of the fp16-from-memory conv replayed as one graph (what the step sees), the time per
launch with the split-K reduction through distributed shared memory (OSM_CONV_SKRED=0) and through the L2 scratch (1, the
default), whether the two outputs are bit-identical, and the per-phase timeline of the LAST launch (median / max over the CTAs; SM clock cycles within a CTA, %globaltimer nanoseconds across CTAs).
"""
import ctypes as C
import math
import os
import statistics
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
TRACE_DIR = os.path.join(HERE, "_trace")
TRACE_LIB = os.path.join(TRACE_DIR, "libosmosis_b200_trace.so")

EVENTS = ["entry", "prologue done", "first ring fill issued", "last TMA issued", "first stage full", "last MMA committed", "accumulator ready",
          "partial staged", "cluster sync 1", "reduced + stored", "cluster sync 2 (DSMEM)", "exit"]

DEFAULT = [(1, 8, 8, 1024, 1024, 9), (1, 8, 8, 1024, 1024, 1), (1, 8, 8, 1024, 3072, 1), (1, 16, 16, 1024, 1024, 9), (1, 16, 16, 1024, 1024, 1),
           (1, 32, 32, 512, 512, 9), (1, 32, 32, 512, 1536, 1), (1, 8, 8, 2048, 1024, 9)]


def build():
    from osmosis_diffusion_code_b200 import build as b
    b.build(force=False, verbose=True)
    os.makedirs(TRACE_DIR, exist_ok=True)
    obj = os.path.join(TRACE_DIR, "conv_tc_trace.o")
    cmd = [b.NVCC, *b.FLAGS, "-DOSM_TRACE=1", "-c", os.path.join(b.CSRC, "conv_tc.cu"), "-o", obj]
    subprocess.run(cmd, check=True, capture_output=True)
    objs = [os.path.join(b.OBJ, s.replace(".cu", ".o")) for s in b.SOURCES if s != "conv_tc.cu"] + [obj]
    subprocess.run([b.NVCC, "-shared", "-o", TRACE_LIB, *objs, "-cudart", "static", "-Xlinker", "--no-undefined"], check=True)
    print("built", TRACE_LIB)


def run(lib, L_, torch, B, H, W, cin, cout, taps, reps=30):
    dev = "cuda"
    k = 3 if taps == 9 else 1
    w = (torch.randn(cout, cin, k, k) / math.sqrt(cin * taps)).to(dev)
    wf = torch.zeros(taps * cout * cin // 2, device=dev); wd = torch.zeros_like(wf)   # fp16 packs in fp32-typed storage
    L_.check(lib.osm_dbg_pack_conv_weight_f16(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, L_.stream()))
    x = torch.randn(B, H, W, cin, device=dev).half(); bias = torch.randn(cout, device=dev); out = torch.empty(B, H, W, cout, device=dev)
    f = lambda: L_.check(lib.osm_dbg_conv_f16(C.c_void_p(x.data_ptr()), cin, L_.ptr(wf), L_.ptr(bias), None, 0, 0, L_.ptr(out), cout, 0, B, H, W,
                                              cin, cout, taps, L_.stream()))
    res = {}
    outs = {}
    for skred in (0, 1):
        os.environ["OSM_CONV_SKRED"] = str(skred)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        outs[skred] = out.clone()
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
        res[skred] = best
        lib.osm_dbg_trace_clear()
        g.replay(); torch.cuda.synchronize()
        n = 256 * 16 * 2
        buf = (C.c_ulonglong * n)()
        lib.osm_dbg_trace_read(buf, n)
        ctas = []
        for c in range(256):
            row = [(buf[(c * 16 + e) * 2], buf[(c * 16 + e) * 2 + 1]) for e in range(12)]
            if row[0][0] and row[11][0]:
                ctas.append(row)
        print(f"{(B, H, W, cin, cout, taps)} skred={skred}: {best:.2f} us per launch in a graph; {len(ctas)} CTAs traced")
        if not ctas or row is None:
            continue
        t0 = min(r[0][1] for r in ctas)
        print(f"    kernel span by %globaltimer (first entry -> last exit): {max(r[11][1] for r in ctas) - t0} ns; "
              f"entry skew {max(r[0][1] for r in ctas) - t0} ns")
        for e in range(1, 12):
            d = [(r[e][0] - r[0][0]) for r in ctas if r[e][0]]
            if d:
                print(f"    {EVENTS[e]:22s} +{statistics.median(d):8.0f} clk (median)  {min(d):8.0f} min  {max(d):8.0f} max   since entry")
    same = torch.equal(outs[0], outs[1])
    print(f"    outputs of the two reductions bit-identical: {same};  {res[0]:.2f} -> {res[1]:.2f} us")


SWEEP = [(1, 8, 8, 1024, 1024, 9), (1, 8, 8, 2048, 1024, 9), (1, 16, 16, 1024, 1024, 9), (1, 16, 16, 2048, 1024, 9), (1, 32, 32, 512, 512, 9),
         (1, 32, 32, 1024, 512, 9), (1, 32, 32, 1024, 1024, 9), (1, 64, 64, 512, 512, 9), (1, 16, 16, 1024, 3072, 1), (1, 8, 8, 1024, 3072, 1),
         (1, 32, 32, 512, 1536, 1), (1, 8, 8, 1024, 1024, 1), (1, 16, 16, 1024, 1024, 1), (1, 32, 32, 512, 512, 1), (1, 64, 64, 512, 512, 1)]


def sweep(lib, L_, torch, B, H, W, cin, cout, taps, reps=32):
    """In-graph time per launch for every forced (BN, split), the launches cycling through enough weight copies to exceed the L2
    (what a step sees: 2.2 GB of weights per pass), fp16 operands from memory."""
    dev = "cuda"
    k = 3 if taps == 9 else 1
    w = (torch.randn(cout, cin, k, k) / math.sqrt(cin * taps)).to(dev)
    n_w = max(2, min(reps, int(200e6 // (taps * cout * cin * 2)) + 1))
    wfs = []
    for _ in range(n_w):
        wf = torch.zeros(taps * cout * cin // 2, device=dev); wd = torch.zeros_like(wf)
        L_.check(lib.osm_dbg_pack_conv_weight_f16(L_.ptr(w), L_.ptr(wf), L_.ptr(wd), cout, cin, cout, cin, taps, L_.stream()))
        wfs.append(wf)
    x = torch.randn(B, H, W, cin, device=dev).half(); bias = torch.randn(cout, device=dev); out = torch.empty(B, H, W, cout, device=dev)

    def t_of():
        try:
            f = lambda i: L_.check(lib.osm_dbg_conv_f16(C.c_void_p(x.data_ptr()), cin, L_.ptr(wfs[i % n_w]), L_.ptr(bias), None, 0, 0, L_.ptr(out),
                                                         cout, 0, B, H, W, cin, cout, taps, L_.stream()))
            f(0); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(reps):
                    f(i)
            g.replay(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / reps * 1e3)
            return best
        except Exception as ex:   # a forced variant the kernel cannot launch (cluster does not fit)
            torch.cuda.synchronize()
            return None
    os.environ.pop("OSM_CONV_FORCE", None)
    os.environ["OSM_CONV_VERBOSE"] = "1"
    base = t_of()
    os.environ.pop("OSM_CONV_VERBOSE", None)
    res = []
    for bn in (64, 128, 256):
        if cout % bn:
            continue
        for sp in (1, 2, 4, 8):
            os.environ["OSM_CONV_FORCE"] = f"{bn},{sp}"
            t = t_of()
            if t is not None:
                res.append((t, bn, sp))
    os.environ.pop("OSM_CONV_FORCE", None)
    print(f"{(B, H, W, cin, cout, taps)} policy {base:6.2f} us | " + "  ".join(f"({bn},{sp}) {t:5.2f}" for t, bn, sp in sorted(res, key=lambda r: (r[1], r[2]))),
          flush=True)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
        sys.exit(0)
    import torch
    from osmosis_diffusion_code_b200 import lib as L_
    L_.LIB_PATH = TRACE_LIB
    lib = L_.load()
    lib.osm_dbg_trace_read.restype = C.c_int
    lib.osm_dbg_trace_read.argtypes = [C.c_void_p, C.c_int]
    lib.osm_dbg_trace_clear.restype = C.c_int
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    for a in sys.argv[1:]:
        if a.startswith("--flags="):   # experiments: 1 = no A loads, 2 = no weight loads, 3 = neither (results are garbage, timing only)
            lib.osm_dbg_trace_flags.argtypes = [C.c_int]
            lib.osm_dbg_trace_flags(int(a.split("=")[1]))
            print("trace flags", a.split("=")[1])
    if "--sweep" in sys.argv:
        for sh in [tuple(int(v) for v in a.split(",")) for a in args] or SWEEP:
            sweep(lib, L_, torch, *sh)
        sys.exit(0)
    shapes = [tuple(int(v) for v in a.split(",")) for a in args] or DEFAULT
    for sh in shapes:
        run(lib, L_, torch, *sh)
