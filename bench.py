#!/usr/bin/env python
"""Headline benchmark: 256x256 RGBD images/sec for a 1000-step guided sampling run (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W               # native sm_100a path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W       # the UNMODIFIED reference (baseline/_ref) on the host CPU
    python bench.py --impl reference-gpu --steps K --warmup W   # the same unmodified reference in eager CUDA on cuda:0

A "step" is ONE guided reverse step of the osmosis_sample_config chain over the whole batch: UNet forward, posterior,
the phi-optimisation / guidance kernel, UNet input-VJP, clamped update + noise.  The chain is the config's sampler
respaced to W+K steps (same per-step work and the same 30 % frozen-phi / 70 % optimised-phi mix as the 1000-step chain),
so      value = images in flight / (1000 * mean step time)      is the 1000-step images/sec, extrapolated from K steps
(K = 1000 makes it exact).  Inputs are synthetic (no network): denoiser-like random weights of the config's
architecture and a physically consistent measurement (osmosis_diffusion_code_b200/synthetic.py).

  value : device-resident - inputs already in HBM when the timed region starts; CUDA events, max over ranks.
  e2e   : the same chain through the public API (sampler.p_sample_loop) with the measurement in pinned HOST memory,
          re-uploaded every step, and the step's loss read back to the host every step (as the reference's progress bar
          does; the read of step k completes while step k+1 runs), plus the final pred_xstart device->host copy.  W untimed
          warm-up steps go through the same API first (graph capture happens there, once per sampler).
  configs : the other BASELINE.json configurations, device-resident, in the same line: config 3 (batch 32 on one GPU),
          config 4 (simulation config, 32 images per GPU = batch 256 on 8 GPUs, finished by an NCCL all-gather of the samples
          and a PSNR of two restored images against the unmodified reference run in eager CUDA on the same GPU) and config 5
          (haze config, 32 images per GPU = batch 128 on 4 GPUs, 250-step respacing, de-gamma'd input).
  gpu_eager_reference : the unmodified reference timed in eager CUDA (cuDNN TF32) on the same GPU, batch 1 - the
          "same algorithm, stock kernels" bar.
All arms print ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "256x256 RGBD images/sec (1000-step guided sampling)"
UNIT = "images/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def measure_tf32_peak(dev, seconds=1.5):
    """cuBLAS TF32 GEMM 8192^3 on this GPU, the same way MEASURED_PEAKS.json measures bf16: best single launch (burst) and
    a back-to-back loop (sustained).  MEASURED_PEAKS.json has no TF32 entry and the dominant kernel computes in TF32."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 0
        t0 = time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(10):
                a @ b
            reps += 10
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        sustained = reps * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        return dict(tf32_tflops=round(best, 1), tf32_tflops_sustained=round(sustained, 1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


def build_native(args, dev, B):
    from tools.profile_step import build
    return build(args.config, B, args.size, conv_mode="tc", dev=dev)


def respaced(a, n_steps):
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import create_sampler
    d = dict(a.diffusion); d["timestep_respacing"] = n_steps
    return create_sampler(**d)


def workload_name(config, B, size, chain_steps, K, T):
    """The same wording in every arm, so that the driver can see that the arms ran the same workload."""
    return (f"{os.path.basename(config)}, batch {B} per GPU, {size}x{size} RGBD, {chain_steps}-step guided chain timed on {K} "
            f"consecutive steps of a {T}-step respacing")


def _phis(opc, k, dflt):
    return [float(v) for v in str(opc.get(k, dflt)).split(",")]


def case_objects(a, dev, B, size, first_image=0, stride=1):
    """operator / conditioning / measurement of `B` synthetic scenes (first_image, first_image + stride, ...) for config `a`."""
    from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator, get_noise
    from osmosis_diffusion_code_b200.guided_diffusion.condition_methods import get_conditioning_method
    from osmosis_diffusion_code_b200.osmosis_utils import data as datao
    from osmosis_diffusion_code_b200.synthetic import synth_measurement
    opc = dict(a.measurement["operator"]); opc["batch_size"] = B
    op = get_operator(device=dev, **opc)
    cond = get_conditioning_method(a.conditioning["method"], op, get_noise(**a.measurement["noise"]), **a.conditioning["params"],
                                   **a.sample_pattern, **a.aux_loss)
    pa, pb = (_phis(opc, "phi_a", "1"), _phis(opc, "phi_b", "1")) if "phi_a" in opc else (_phis(opc, "phi_ab", "1"),) * 2
    y = torch.cat([synth_measurement(first_image + i * stride, size, pa, pb, _phis(opc, "phi_inf", "0.2,0.4,0.7"),
                                     depth_type=opc.get("depth_type"))[0] for i in range(B)], 0).to(dev)
    if getattr(a, "degamma_input", False):
        y = datao.degamma_input(y)                  # osmosis_sampling.py:173-175
    return op, cond, y


def time_case(model, cfg_path, dev, B, K, W, size, rank, world, local, chain_steps=None, cuda_graph=True, keep=False):
    """Device-resident timing of one BASELINE configuration: W warm-up + K timed guided steps of the config's chain respaced
    to W + K steps, B images per GPU; max over ranks.  Returns (row dict, objects kept for the caller when keep)."""
    import torch.distributed as dist
    from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import FusedStepper
    a = arguments_from_file(cfg_path)
    base_T = int(a.diffusion["steps"])
    chain = int(chain_steps or a.diffusion.get("timestep_respacing") or base_T)
    # images of rank r: r, r + world, ... (sharding.shard_indices)
    op, cond, y = case_objects(a, dev, B, size, first_image=rank, stride=world)
    sampler = respaced(a, min(K + W, base_T))
    T = sampler.num_timesteps
    torch.manual_seed(a.manual_seed)
    img = torch.randn(B, 4, size, size, device=dev)
    stepper = FusedStepper(sampler, model, cond, img, y, a.sample_pattern, cuda_graph=cuda_graph)
    idxs = [T - 1 - (i % T) for i in range(W + K)]
    for idx in idxs[:W]:
        stepper.step(idx)
    clocks = ClockSampler(local) if rank == 0 else None
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    if clocks: clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for idx in idxs[W:]:
        stepper.step(idx)
    e1.record()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms[0]) / K
    clk = clocks.stop() if clocks else None
    row = dict(config=os.path.basename(cfg_path), batch_per_gpu=B, global_batch=B * world, chain_steps=chain,
               workload=workload_name(cfg_path, B, size, chain, K, T), ms_per_step=ms_step,
               ms_per_image_step=ms_step / B, images_per_s=world * B / (chain * ms_step / 1e3), clocks=clk,
               finite=bool(torch.isfinite(img).all()), workspace_gb=round(model.workspace_bytes(B, size, size) / 1e9, 2))
    return row, (dict(a=a, op=op, cond=cond, y=y, stepper=stepper, img=img) if keep else None)


def gather_finished(kept, B, world, dev):
    """The one optional collective of the path: all-gather of the finished samples, phi and losses over NCCL.
    Returns its device time in ms (max over ranks) and the gathered shape."""
    import torch.distributed as dist
    from osmosis_diffusion_code_b200 import sharding
    st = kept["stepper"].st
    x0, phi, loss = st["x0"].contiguous(), kept["op"].phi.detach().contiguous(), st["losses"][:, 0].contiguous()
    n = B * world
    sharding.gather_images(loss, n)        # warm the communicator
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gx0 = sharding.gather_images(x0, n); gphi = sharding.gather_images(phi, n); gl = sharding.gather_images(loss, n)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = bool(torch.equal(gx0[dist.get_rank()::world], x0)) and tuple(gphi.shape) == (n, 9) and gl.numel() == n
    return dict(ms=float(t[0]), bytes_per_rank=int(x0.numel() * 4 + phi.numel() * 4 + loss.numel() * 4),
                gathered=list(gx0.shape), own_rows_intact=ok, backend="nccl")


def psnr_vs_reference_gpu(model, cfg_path, dev, size, n_images, T, cuda_graph=True):
    """BASELINE config 4's check: final pred_xstart[:, :3] of `n_images` scenes sampled here as ONE batch (shared x_T and
    step noise, i.e. what the reference's per-image reseed gives) against the unmodified reference run per image in eager
    CUDA on this GPU (cuDNN TF32), both free-running from `manual_seed` over the same T-step respaced chain.  Data range 2."""
    from baseline import ref_driver as rd
    from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file
    a = arguments_from_file(cfg_path)
    op, cond, y = case_objects(a, dev, n_images, size)
    sampler = respaced(a, T)
    torch.manual_seed(a.manual_seed)
    x_start = torch.randn(1, 4, size, size, device=dev).expand(n_images, -1, -1, -1).contiguous()
    _img, _vd, _loss, x0 = sampler.p_sample_loop(model=model, x_start=x_start, measurement=y, measurement_cond_fn=cond.conditioning,
                                                 record=False, save_root=None, pretrain_model="osmosis", rgb_guidance=False,
                                                 sample_pattern=a.sample_pattern, noise_mode="shared", cuda_graph=cuda_graph)
    R = rd.import_reference()
    rargs = R.utils.arguments_from_file(cfg_path)
    rargs.diffusion = dict(rargs.diffusion); rargs.diffusion["timestep_respacing"] = T
    rmodel = rd.reference_model(rargs, dev)
    out = []
    import contextlib
    for i in range(n_images):
        roperator, rcond, rsampler = rd.reference_pieces(rargs, dev, batch=1)
        torch.manual_seed(rargs.manual_seed)
        xs = torch.randn([1, 4, size, size], device=dev).requires_grad_()
        with contextlib.redirect_stderr(open(os.devnull, "w")):
            _i, _v, _l, rx0 = rsampler.p_sample_loop(model=rmodel, x_start=xs, measurement=y[i:i + 1], measurement_cond_fn=rcond.conditioning,
                                                     pretrain_model="osmosis", rgb_guidance=False, sample_pattern=rargs.sample_pattern,
                                                     record=False, save_root=None, image_idx=i, record_every=200,
                                                     original_file_name="bench", save_grids_path=None, global_iteration=0)
        mse = float(((x0[i, :3].double() - rx0[0, :3].double()) ** 2).mean())
        out.append(round(10 * math.log10(4.0 / max(mse, 1e-30)), 2))
    del rmodel
    torch.cuda.empty_cache()
    return dict(psnr_db=out, images=n_images, chain=f"{T}-step respacing of {os.path.basename(cfg_path)}, free-running from manual_seed",
                against="unmodified reference (baseline/_ref) in eager CUDA on the same GPU, cuDNN TF32 convolutions", data_range=2.0)


def gpu_eager_reference(args, K, W):
    from baseline import ref_driver as rd
    r = rd.time_reference_chain(args.config, "cuda", steps=K, warmup=W, size=args.size)
    t = sum(r["step_s"]) / len(r["step_s"])
    torch.cuda.empty_cache()
    return dict(ms_per_step=t * 1e3, images_per_s=1.0 / (1000.0 * t), steps=len(r["step_s"]), warmup=W, batch=1,
                how="unmodified reference (baseline/_ref) p_sample_loop in eager CUDA on the same GPU, cuDNN default (TF32 convs), "
                    "host clock between UNet calls after a device sync", finite=r["finite"])


def run_native(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, args.warmup
    graph = not args.no_cuda_graph
    a, model, op, cond, _, y = build_native(args, dev, B)
    base_T = int(a.diffusion["steps"])
    sampler = respaced(a, min(K + W, base_T))    # K + W > 1000: the full chain, indices wrap around
    T = sampler.num_timesteps
    torch.manual_seed(a.manual_seed)
    img = torch.randn(B, 4, args.size, args.size, device=dev)
    from osmosis_diffusion_code_b200.guided_diffusion.gaussian_diffusion import FusedStepper
    stepper = FusedStepper(sampler, model, cond, img, y, a.sample_pattern, cuda_graph=graph)
    step = stepper.step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    idxs = [T - 1 - (i % T) for i in range(W + K)]
    for idx in idxs[:W]:
        step(idx)
    # ---- timed region (device-resident) ----
    clocks = ClockSampler(local) if rank == 0 else None
    barrier()
    if clocks: clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profiler_range:   # ncu --profile-from-start off: the launch list then holds the timed steps only
        torch.cuda.profiler.start()
    e0.record()
    for idx in idxs[W:]:
        step(idx)
    e1.record()
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks else None
    finite = bool(torch.isfinite(img).all())
    # ---- e2e through the public API, host buffers ----
    from osmosis_diffusion_code_b200.guided_diffusion.measurements import get_operator, get_noise
    from osmosis_diffusion_code_b200.guided_diffusion.condition_methods import get_conditioning_method
    opc = dict(a.measurement["operator"]); opc["batch_size"] = B
    op2 = get_operator(device=dev, **opc)
    cond2 = get_conditioning_method(a.conditioning["method"], op2, get_noise(**a.measurement["noise"]), **a.conditioning["params"],
                                    **a.sample_pattern, **a.aux_loss)
    Te = min(K, base_T)
    samp2 = respaced(a, Te)
    y_host = y.cpu().pin_memory()
    torch.manual_seed(a.manual_seed)
    x_start = torch.randn(B, 4, args.size, args.size, device=dev)
    host_loss = []
    loop_kw = dict(model=model, measurement=y_host, measurement_cond_fn=cond2.conditioning, record=False, save_root=None,
                   pretrain_model="osmosis", rgb_guidance=False, sample_pattern=a.sample_pattern, cuda_graph=graph)
    # untimed warm-up through the same API (W steps): the sampler keeps its device state and the captured step graph
    # between p_sample_loop calls, as it does between the images of a sampling run
    samp2.p_sample_loop(x_start=x_start, max_steps=min(W, Te), progress=lambda idx, loss: None, **loop_kw)
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    left = K
    while left > 0:                                  # K > 1000: several full chains
        n = min(left, Te)
        _img, _vd, _loss, x0_cpu = samp2.p_sample_loop(x_start=x_start, max_steps=n, progress=lambda idx, loss: host_loss.append(loss),
                                                       **loop_kw)
        left -= n
    f1.record()
    torch.cuda.synchronize()
    e2e_s = max(time.perf_counter() - t0, f0.elapsed_time(f1) / 1e3)  # device events and host clock agree; keep the larger
    h2d = y_host.numel() * 4
    d2h = B * 4 * 4 + x0_cpu.numel() * 4 // K
    # ---- reductions over ranks: max time ----
    times = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(times[0]), float(times[1])
    ms_step = ms / K
    value = world * B / (1000.0 * ms_step / 1e3)
    e2e_value = world * B / (1000.0 * (e2e_ms / K) / 1e3)
    roofline, fl, bl = None, 0, 0
    if rank == 0:
        roofline, fl, bl = dominant_roofline(args, model, dev)
    # ---- the other BASELINE configurations (device-resident, same K / W), every rank takes part ----
    table, gather, psnr, eager, errors = [], None, None, None, {}
    del stepper, samp2, cond2, op2
    if not args.no_table:
        TB = args.table_batch
        cfgd = os.path.join(ROOT, "configs")
        cases = [("config 3: batch 32 on one GPU", args.config, None), ("config 4: simulation, 32 images / GPU (batch 256 on 8 GPUs)",
                 os.path.join(cfgd, "osmosis_simulation_sample_config.yaml"), None),
                 ("config 5: haze, 32 images / GPU (batch 128 on 4 GPUs), 250-step respacing, de-gamma'd input",
                  os.path.join(cfgd, "osmosis_haze_sample_config.yaml"), 250)]
        for name, cpath, chain in cases:
            try:
                is_sim = "simulation" in cpath
                row, kept = time_case(model, cpath, dev, TB, K, W, args.size, rank, world, local, chain_steps=chain, cuda_graph=graph,
                                      keep=is_sim and world > 1)
                row["baseline_config"] = name
                if kept is not None:
                    gather = gather_finished(kept, TB, world, dev)
                    del kept
                table.append(row)
            except Exception as e:   # a table row must not cost the headline number
                errors[name] = f"{type(e).__name__}: {e}"
                barrier()
    if rank == 0:
        from baseline import ref_driver as rd
        if rd.available() and not args.no_eager_reference:
            try:
                psnr = psnr_vs_reference_gpu(model, os.path.join(ROOT, "configs", "osmosis_simulation_sample_config.yaml"), dev, args.size,
                                             2, min(K + W, base_T), cuda_graph=graph)
            except Exception as e:
                errors["psnr"] = f"{type(e).__name__}: {e}"
            try:
                eager = gpu_eager_reference(args, K, W)
            except Exception as e:
                errors["gpu_eager_reference"] = f"{type(e).__name__}: {e}"
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_step, higher_is_better=True,
               scaling="weak", vs_baseline=None, dtype="f16|tf32 operands, f32 accumulate", data="synthetic",
               dtype_note="tensor-core operands carry an 11-bit significand everywhere (fp16 where a kernel of ours produces or converts the "
                          "operand, TF32 elsewhere: what cuDNN runs for the reference on a GPU); accumulation, GroupNorm, softmax, sampler "
                          "and guidance are fp32; OSM_CONV_F16=0 selects TF32 for every conv",
               config=dict(workload=workload_name(args.config, B, args.size, 1000, K, T), batch_per_gpu=B,
                           global_batch=B * world, image=args.size, unet_params=model.num_params(), parallelism=f"dp{world} (batch-sharded, no collective)",
                           l2="per-step working set (2.2 GB weights + 1.6 GB activations per image) exceeds the 126 MB L2"),
               clocks=clk, finite=finite,
               e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / K,
                        api="sampler.p_sample_loop(measurement=<pinned host tensor>, progress=<host callback>)"),
               gpu_launches=K * (fl + bl + 4), roofline=roofline, configs=table)
    if gather is not None:
        out["final_gather"] = gather
    if psnr is not None:
        out["psnr_vs_reference"] = psnr
    if eager is not None:
        eager["speedup_of_native_b1"] = round(eager["ms_per_step"] / ms_step, 2) if B == 1 else None
        out["gpu_eager_reference"] = eager
    if errors:
        out["errors"] = errors
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, steps=4, warmup=1, budget_s=60.0)
    print(json.dumps(out))
    if world > 1: dist.destroy_process_group()


def dominant_roofline(args, model, dev):
    """Roofline of the dominant kernel: the 3x3 conv 256->256 @ 256x256 (tcgen05 implicit GEMM), timed per op inside the step."""
    peaks = _peaks()
    fl, bl = model.launch_counts()
    prof = model.profile_ops(0) + model.profile_ops(1)
    conv = [o for o in prof if o["kind"] == "conv"]
    dom = [o for o in conv if o["dims"][:5] == [args.size, args.size, 256, 256, 9]] or conv
    dom_ms = sum(o["ms"] for o in dom) / len(dom)
    dom_tf = dom[0]["flops"] / dom_ms / 1e9
    opk = sorted({o["dims"][5] >> 1 for o in dom})      # operand type of those launches: 0 TF32, 1 fp16 halo kernel, 2 fp16 from memory
    f16 = opk == [1]
    by_operand = {}
    for o in conv:
        d = by_operand.setdefault(("tf32", "fp16 halo kernel", "fp16 from memory")[o["dims"][5] >> 1], [0.0, 0.0, 0])
        d[0] += o["flops"]; d[1] += o["ms"]; d[2] += 1
    conv_all_tf = sum(o["flops"] for o in conv) / sum(o["ms"] for o in conv) / 1e9
    norm = [o for o in prof if o["kind"].startswith("gn")]
    norm_gbs = sum(o["bytes"] for o in norm) / sum(o["ms"] for o in norm) / 1e6
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_conv_traffic.json")
    tsrc = None
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        ent = tj.get("fp16_halo" if f16 else "tf32_pair", tj)
        traffic, tsrc = ent.get("dram_bytes_per_launch"), ent.get("source")
    tf32 = measure_tf32_peak(dev)
    kname = ("conv_tc_halo16_2sm_kernel (3x3 256->256 @256x256: tcgen05 cta_group::2 kind::f16 implicit GEMM on halo tiles, fp32 input "
             "converted / GroupNorm+SiLU-transformed to fp16 in shared memory, fp32 accumulation in TMEM; fwd + dgrad launches)") if f16 else \
            "conv_tc_persist_2sm_kernel<6,.> (3x3 256->256 @256x256 tcgen05 cta_group::2 TF32 implicit GEMM, fwd + dgrad launches)"
    roofline = dict(bound="tensor", kernel=kname,
                    achieved=round(dom_tf, 1), peak=peaks["bf16_sustained"], unit="TFLOP/s",
                    frac=round(dom_tf / peaks["bf16_sustained"], 4), traffic=traffic,
                    traffic_source=tsrc or "ncu --set full capture of this kernel (profiles/ncu_conv_traffic.json), per launch",
                    peak_source=f"{peaks['source']} bf16 sustained from MEASURED_PEAKS.json (kernel timed inside the step)",
                    dtype_note=("fp16 operands (11-bit significand, the same as TF32 = the reference's own GPU arithmetic for conv), fp32 accumulation; "
                                "the dense 16-bit peak is the denominator" if f16 else
                                "the kernel computes in TF32 (the reference's own GPU arithmetic for conv); cuBLAS TF32 8192^3 measured "
                                "in this run the same way is the like-for-like denominator"),
                    convs_by_operand={k: dict(tflops=round(v[0] / v[1] / 1e9, 1), ms=round(v[1], 3), launches=v[2]) for k, v in by_operand.items()},
                    tf32_peak_measured=tf32, frac_of_tf32_sustained=round(dom_tf / tf32["tf32_tflops_sustained"], 4),
                    flops_per_launch=dom[0]["flops"], launches_averaged=len(dom), ms_per_launch=round(dom_ms, 4),
                    all_convs_tflops=round(conv_all_tf, 1), groupnorm_kernels_gbs=round(norm_gbs, 0), hbm_peak_gbs=peaks["hbm"],
                    step_ms_by_kind={k: round(sum(o["ms"] for o in prof if o["kind"] == k), 3) for k in sorted({o["kind"] for o in prof})})
    # secondary rooflines (informational): the HBM-bound GroupNorm apply kernel and the fused tcgen05 attention forward
    def _avg(kind, dims3):
        sel = [o for o in prof if o["kind"] == kind and o["dims"][:3] == dims3]
        return (sum(o["ms"] for o in sel) / len(sel), sel[0]) if sel else (None, None)
    extra = {}
    gms, g0 = _avg("gn_apply", [args.size, args.size, 256])
    if gms:
        extra["groupnorm"] = dict(bound="hbm", kernel=f"gn_apply_kernel (GroupNorm32 + scale-shift + SiLU, {args.size}x{args.size}x256, read + write)",
                                  achieved=round(g0["bytes"] / gms / 1e6, 1), peak=peaks["hbm"], unit="GB/s",
                                  frac=round(g0["bytes"] / gms / 1e6 / peaks["hbm"], 4), bytes_per_launch=g0["bytes"], ms_per_launch=round(gms, 4))
    ams, a0 = _avg("attn_fwd", [(args.size // 8) ** 2, 512, 8])
    if ams:
        extra["attention"] = dict(bound="latency (tensor pipe far from saturated: 0.5 % of the step's FLOPs)",
                                  kernel="tok_to_chan_kernel + flash_fwd_kernel (tcgen05 kind::tf32, L=1024, 8 heads x 64 ch)",
                                  achieved=round(a0["flops"] / ams / 1e9, 1), peak=peaks["bf16_sustained"], unit="TFLOP/s",
                                  frac=round(a0["flops"] / ams / 1e9 / peaks["bf16_sustained"], 4), ms_per_launch=round(ams, 4))
    roofline["other_kernels"] = extra
    return roofline, fl, bl


# ----------------------------------------------------------------------------------- the reference arms (CPU / eager CUDA)


def _reference_chain(args, device, K, W, budget_s):
    """Per-step times of the unmodified reference (baseline/_ref) on `device`; falls back to the oracle port (kind 'port')
    only when the reference copy is absent."""
    from baseline import ref_driver as rd
    cores = os.cpu_count() or 1
    if rd.available():
        r = rd.time_reference_chain(args.config, device, steps=K, warmup=W, size=args.size, budget_s=budget_s,
                                    threads=cores if device == "cpu" else None)
        return r["step_s"], r["T"], "reference", r["finite"], cores
    if device != "cpu":
        raise SystemExit("baseline/_ref is missing: the eager-CUDA reference arm needs the reference copy")
    from osmosis_diffusion_code_b200.osmosis_utils.utils import arguments_from_file
    a = arguments_from_file(args.config)
    a.diffusion = dict(a.diffusion); a.diffusion["timestep_respacing"] = K + W
    step, T, cores = cpu_step_runner(args, a)
    idxs = list(range(T))[::-1]
    t_begin = time.perf_counter()
    for idx in idxs[:W]:
        step(idx)
    ts = []
    for idx in idxs[W:]:
        ts.append(step(idx))
        if time.perf_counter() - t_begin > budget_s and len(ts) >= 2:
            break
    return ts, T, "port", None, cores


def cpu_step_runner(args, a):
    """Fallback only (no baseline/_ref): one guided step of the oracle port on the host CPU."""
    from oracle import osmosis_oracle as orc
    from osmosis_diffusion_code_b200.synthetic import synth_state_dict, synth_measurement
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = orc.UNetConfig.from_create_model_kwargs(**a.unet_model)
    sd = synth_state_dict(list(orc.param_shapes(cfg).items()), cfg.model_channels, seed=7, delta=0.05)
    ydict = dict(vars(a))
    tab, ospec, gspec, phis, names = orc.specs_from_config(ydict, 1)
    o = a.measurement["operator"]
    f = lambda k: [float(v) for v in str(o[k]).split(",")]
    pa, pb = (f("phi_a"), f("phi_b")) if "phi_a" in o else (f("phi_ab"), f("phi_ab"))
    y, _ = synth_measurement(0, args.size, pa, pb, f("phi_inf"), depth_type=o.get("depth_type"))
    torch.manual_seed(a.manual_seed)
    state = dict(x=torch.randn(1, 4, args.size, args.size), phis=phis)

    def step(idx):
        t0 = time.perf_counter()
        r = orc.guided_step(sd, cfg, tab, ospec, gspec, state["x"], y, state["phis"], idx, torch.randn(1, 4, args.size, args.size))
        state["x"], state["phis"] = r["x_next"], r["phis"]
        return time.perf_counter() - t0

    return step, tab.num_timesteps, cores


def cpu_baseline(args, steps, warmup, budget_s):
    """The reference's own CPU path on this box's host cores, a bounded sample: `warmup` + `steps` steps of the chain."""
    ts, T, kind, finite, cores = _reference_chain(args, "cpu", steps, warmup, budget_s)
    t = statistics.median(ts)
    what = "the unmodified reference (baseline/_ref) through its own p_sample_loop" if kind == "reference" else "oracle port of the reference"
    return dict(value=1.0 / (1000.0 * t), unit=UNIT, cores=cores, kind=kind,
                sample=f"{len(ts)} guided steps (after {warmup} warm-up) of a {T}-step respacing of the same chain at batch 1, {what}, true "
                       f"fp32, torch {torch.__version__} CPU kernels on {cores} threads, median {t:.2f} s/step; images/s extrapolated to "
                       f"1000 steps", seconds_per_step=t, raw_step_seconds=[round(v, 3) for v in ts])


def run_reference(args, device):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    if device == "cuda":
        if not torch.cuda.is_available():
            raise SystemExit("--impl reference-gpu needs a CUDA device")
        torch.cuda.set_device(0)
    ts, T, kind, finite, cores = _reference_chain(args, device, K, W, args.cpu_budget_s if device == "cpu" else None)
    t = sum(ts) / len(ts)
    value = 1.0 / (1000.0 * t)
    if device == "cpu":
        what = (f"{len(ts)} of the requested {K} guided steps after {W} warm-up steps" + (f" (time-capped at {args.cpu_budget_s:.0f} s)" if len(ts) < K else "") +
                f", batch 1, true fp32 on {cores} host threads, " +
                ("the UNMODIFIED reference (baseline/_ref) through its own create_model / get_operator / get_conditioning_method / "
                 "create_sampler / p_sample_loop" if kind == "reference" else "oracle port of the reference (baseline/_ref absent)"))
    else:
        what = (f"{len(ts)} guided steps after {W} warm-up steps, batch 1, the UNMODIFIED reference (baseline/_ref) in eager CUDA on cuda:0 "
                f"({torch.cuda.get_device_name(0)}), cuDNN default TF32 convolutions")
    out = dict(impl="reference" if device == "cpu" else "reference-gpu", metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=len(ts),
               warmup=W, ms_per_step=t * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f32" if device == "cpu" else "tf32 (cuDNN default)", data="synthetic",
               config=dict(workload=workload_name(args.config, 1, args.size, 1000, K, T), batch_per_gpu=1, global_batch=1, image=args.size,
                           note="the reference runs batch 1 only (gaussian_diffusion.py:216); under torchrun rank 0 alone runs it"),
               cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=what), finite=finite,
               e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-gpu"])
    ap.add_argument("--batch", type=int, default=1, help="images per GPU")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--config", default=os.path.join(ROOT, "configs", "osmosis_sample_config.yaml"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true")
    ap.add_argument("--no-table", action="store_true", help="skip the table of the other BASELINE configurations")
    ap.add_argument("--table-batch", type=int, default=32, help="images per GPU of the configs table (BASELINE configs 3-5)")
    ap.add_argument("--no-eager-reference", action="store_true", help="skip the eager-CUDA reference timing and the PSNR check")
    ap.add_argument("--profiler-range", action="store_true", help="cudaProfilerStart/Stop around the timed device-resident region")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, "cpu")
    elif args.impl == "reference-gpu":
        run_reference(args, "cuda")
    else:
        run_native(args)


if __name__ == "__main__":
    main()
