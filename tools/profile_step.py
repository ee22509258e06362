"""Per-op breakdown of one guided step at full size (development / profiles tool, not a benchmark).

    python tools/profile_step.py --batch 1 [--size 256] [--config configs/osmosis_sample_config.yaml]

Builds the config's UNet with synthetic weights, runs a few guided steps, then one event-timed pass of the forward and
input-VJP programs (osm_unet_profile_ops) and prints the time grouped by kernel class and shape with achieved
TFLOP/s / GB/s.  Writes the table to --out as JSON when given.
"""
import argparse
import collections
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file  # noqa: E402
from osmosis_diffusion_code_b200.guided_diffusion.unet import create_model  # noqa: E402
from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import create_sampler  # noqa: E402
from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator, get_noise  # noqa: E402
from osmosis_diffusion_code_b200.guided_diffusion.condition_methods import get_conditioning_method  # noqa: E402
from osmosis_diffusion_code_b200.synthetic import synth_state_dict, synth_measurement  # noqa: E402


def build(cfg_path, B, size, conv_mode="tc", dev="cuda"):
    a = arguments_from_file(cfg_path)
    um = dict(a.unet_model); um["model_path"] = ""
    t0 = time.time()
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # create_model reports the missing checkpoint on stdout, like the reference
        model = create_model(**um, conv_mode=conv_mode)
    sd = synth_state_dict(model.param_specs(), um["num_channels"], seed=7, delta=0.05)
    model.load_state_dict(sd); del sd
    model.to(dev)
    opc = dict(a.measurement["operator"]); opc["batch_size"] = B
    op = get_operator(device=dev, **opc)
    cond = get_conditioning_method(a.conditioning["method"], op, get_noise(**a.measurement["noise"]), **a.conditioning["params"],
                                   **a.sample_pattern, **a.aux_loss)
    sampler = create_sampler(**a.diffusion)
    ph = lambda k, d: [float(v) for v in str(opc.get(k, d)).split(",")]
    if "phi_a" in opc:
        pa, pb = ph("phi_a", "1"), ph("phi_b", "1")
    else:
        pa = pb = ph("phi_ab", "1")
    ys = [synth_measurement(i, size, pa, pb, ph("phi_inf", "0.2,0.4,0.7"), depth_type=opc.get("depth_type"))[0] for i in range(B)]
    y = torch.cat(ys, 0).to(dev)
    print(f"# built model ({model.num_params()/1e6:.1f} M params) + inputs in {time.time()-t0:.1f} s", file=sys.stderr)
    return a, model, op, cond, sampler, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--config", default="configs/osmosis_sample_config.yaml")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    B = args.batch
    a, model, op, cond, sampler, y = build(args.config, B, args.size)
    torch.manual_seed(a.manual_seed)
    img = torch.randn(B, 4, args.size, args.size, device="cuda")
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import FusedStepper
    stepper = FusedStepper(sampler, model, cond, img, y, a.sample_pattern, cuda_graph=not args.no_graph)
    T = sampler.num_timesteps

    def step(idx, freeze):
        stepper.step(idx, freeze=freeze)

    for k in range(3):
        step(T - 1 - k, True)
    torch.cuda.synchronize()
    res = {}
    for name, freeze, idx0 in (("frozen", True, T - 4), ("optimised", False, int(0.6 * T))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            step(idx0 - k, freeze)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / args.steps
        print(f"step ({name} phase): {res[name]:.2f} ms   finite={bool(torch.isfinite(img).all())}")
    fl, bl = model.launch_counts()
    print(f"launches: fwd {fl} vjp {bl}; fwd algorithmic GFLOP {model.forward_flops()/1e9:.1f}")
    table = []
    for which, nm in ((0, "fwd"), (1, "vjp")):
        ops = model.profile_ops(which)
        agg = collections.OrderedDict()
        for o in ops:
            key = (o["kind"], tuple(o["dims"]))
            d = agg.setdefault(key, dict(n=0, ms=0.0, flops=0.0, bytes=0.0))
            d["n"] += 1; d["ms"] += o["ms"]; d["flops"] += o["flops"]; d["bytes"] += o["bytes"]
        tot = sum(o["ms"] for o in ops)
        print(f"\n== {nm}: {len(ops)} ops, {tot:.2f} ms (event-timed per op)")
        bykind = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
        for (kind, dims), d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            bykind[kind][0] += d["ms"]; bykind[kind][1] += d["flops"]; bykind[kind][2] += d["bytes"]
            row = dict(prog=nm, kind=kind, dims=list(dims), n=d["n"], ms=d["ms"], tflops=d["flops"] / d["ms"] / 1e9 if d["ms"] else 0,
                       gbs=d["bytes"] / d["ms"] / 1e6 if d["ms"] else 0)
            table.append(row)
            if d["ms"] > 0.01 * tot:
                print(f"  {kind:9s} {str(list(dims)):34s} x{d['n']:<3d} {d['ms']:8.3f} ms  {row['tflops']:8.1f} TF/s {row['gbs']:8.0f} GB/s")
        for kind, (ms, f, b) in sorted(bykind.items(), key=lambda kv: -kv[1][0]):
            print(f"  [{kind:9s}] {ms:8.3f} ms ({100*ms/tot:5.1f} %)  {f/ms/1e9 if ms else 0:8.1f} TF/s  {b/ms/1e6 if ms else 0:8.0f} GB/s")
    if args.out:
        json.dump(dict(batch=B, size=args.size, step_ms=res, table=table), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
