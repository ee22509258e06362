"""Samplers: schedule tables, timestep respacing and the guided reverse loop.

Mirrors the reference `guided_diffusion/gaussian_diffusion.py`: the sampler registry and `create_sampler`
(:19-62), `GaussianDiffusion` (:65-365), `space_timesteps` (:373-426), `SpacedDiffusion` / `_WrappedModel`
(:429-489), `DDPM` (:492-503), `get_named_beta_schedule` (:542-566), `extract_and_expand` (:593-597).

`p_sample_loop` keeps the reference's signature and return value `(img, variable_dict, loss, pred_xstart on
CPU)`.  Two execution paths produce the same arithmetic:

  * fused (default when the model is the native UNet and the conditioning method is the native `osmosis`
    one): per step  UNet forward -> posterior kernel -> guidance/phi kernel -> posterior VJP -> UNet input-VJP
    -> sampler-update kernel, all asynchronous on the current stream with no host synchronisation; the step
    index and the freeze flag live in device memory so the launch sequence is identical every step.
  * autograd-compatible (any other model / conditioning callable): the reference's call sequence
    (`p_mean_variance` -> `measurement_cond_fn(...)` -> noise), differentiable through autograd Functions.

Deliberate deviations: batches of B > 1 run as B independent chains instead of crashing
(:216); no per-step `.item()` / `.cpu()` (the tqdm postfix of :276-296) - loss and phi are read once at the end.
The dead `q_sample(measurement)` draw (:241) IS mirrored because it advances the RNG stream.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import lib as _lib
from ..osmosis_utils import utils as utilso
from .posterior_mean_variance import PosteriorFn, get_mean_processor, get_var_processor

__SAMPLER__ = {}


def register_sampler(name: str):
    def wrapper(cls):
        if __SAMPLER__.get(name, None):
            raise NameError(f"Name {name} is already registered!")
        __SAMPLER__[name] = cls
        return cls
    return wrapper


def get_sampler(name: str):
    if __SAMPLER__.get(name, None) is None:
        raise NameError(f"Name {name} is not defined!")
    return __SAMPLER__[name]


def create_sampler(sampler, steps, noise_schedule, model_mean_type, model_var_type, dynamic_threshold, clip_denoised,
                   rescale_timesteps, timestep_respacing="", **kwargs):
    cls = get_sampler(name=sampler)
    betas = get_named_beta_schedule(noise_schedule, steps)
    if not timestep_respacing:
        timestep_respacing = [steps]
    return cls(use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas, model_mean_type=model_mean_type,
               model_var_type=model_var_type, dynamic_threshold=dynamic_threshold, clip_denoised=clip_denoised,
               rescale_timesteps=rescale_timesteps, annealing_time=kwargs.get("annealing_time", False))


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        abar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_diffusion_timesteps
        return np.array([min(1 - abar((i + 1) / n) / abar(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """Which original timesteps a respaced chain keeps: "ddimN", "a,b,c" per-section counts, or an int."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == desired:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    elif isinstance(section_counts, int):
        section_counts = [section_counts]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


def extract_and_expand(array, time, target):
    array = torch.from_numpy(np.asarray(array)).to(target.device)[time].float()
    while array.ndim < target.ndim:
        array = array.unsqueeze(-1)
    return array.expand_as(target)


class GaussianDiffusion:
    def __init__(self, betas, model_mean_type, model_var_type, dynamic_threshold, clip_denoised, rescale_timesteps, **kwargs):
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1, "betas must be 1-D"
        assert (0 < betas).all() and (betas <= 1).all(), "betas must be in (0..1]"
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        self.rescale_timesteps = rescale_timesteps
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.mean_processor = get_mean_processor(model_mean_type, betas=betas, dynamic_threshold=dynamic_threshold,
                                                 clip_denoised=clip_denoised)
        self.var_processor = get_var_processor(model_var_type, betas=betas)

    @property
    def posterior_flags(self):
        """OSM_POST_* flags of this sampler's mean / variance processors and clip_denoised (include/osmosis_b200.h)."""
        return self.mean_processor.flags | self.var_processor.flags

    # ---- pieces of the reference API --------------------------------------------------------------------
    def q_sample(self, x_start, t):
        """sqrt(abar_t) x + sqrt(1 - abar_t) randn - the osmosis path only needs its RNG side effect (:241)."""
        noise = torch.randn_like(x_start)
        c1 = extract_and_expand(self.sqrt_alphas_cumprod, t, x_start)
        c2 = extract_and_expand(self.sqrt_one_minus_alphas_cumprod, t, x_start)
        return c1 * x_start + c2 * noise

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def p_mean_variance(self, model, x, t):
        model_output = model(x, self._scale_timesteps(t))
        if model_output.shape[1] != 2 * x.shape[1]:
            raise NotImplementedError("the native posterior kernel expects a learned-variance model (2C output channels)")
        coef = self.mean_processor.table.on(x.device)
        x0, mean, logvar = PosteriorFn.apply(x, model_output, coef, t.to(torch.int32).contiguous(), self.posterior_flags)
        return {"mean": mean, "variance": torch.exp(logvar.detach()), "log_variance": logvar, "pred_xstart": x0}

    def p_sample(self, model, x, t):
        raise NotImplementedError

    # ---- the loop ---------------------------------------------------------------------------------------
    def _model_timestep(self, idx):
        return float(idx) * (1000.0 / self.num_timesteps) if self.rescale_timesteps else float(idx)

    def p_sample_loop(self, model, x_start, measurement, measurement_cond_fn, record, save_root, pretrain_model=None,
                      image_idx=None, record_every=150, rgb_guidance=False, sample_pattern=None, **kwargs):
        if not x_start.is_cuda:
            raise _lib.OsmError("p_sample_loop needs CUDA tensors (no CPU fallback)")
        from .condition_methods import PosteriorSamplingOsmosis, PosteriorSampling
        from .unet import UNetModel
        cond = getattr(measurement_cond_fn, "__self__", None)
        if pretrain_model != "osmosis" or rgb_guidance:
            # "almost original dps code - rgb_guidance" (gaussian_diffusion.py:232-233, 299-306): p_sample, then the `ps`
            # conditioning; returns only the image (:339-340)
            idxs = list(range(self.num_timesteps))[::-1]
            if kwargs.get("max_steps") is not None:
                idxs = idxs[:kwargs["max_steps"]]
            fused = (isinstance(model, UNetModel) and isinstance(cond, PosteriorSampling)
                     and getattr(measurement_cond_fn, "__func__", None) is PosteriorSampling.conditioning
                     and kwargs.get("fused", True))
            if fused:
                return self._loop_fused_ps(model, cond, x_start, measurement, idxs, kwargs.get("noise_mode", "batch"),
                                           kwargs.get("progress"), kwargs.get("cuda_graph", True))
            return self._loop_autograd_ps(model, measurement_cond_fn, x_start, measurement, idxs)
        fused = (isinstance(model, UNetModel) and isinstance(cond, PosteriorSamplingOsmosis)
                 and getattr(measurement_cond_fn, "__func__", None) is PosteriorSamplingOsmosis.conditioning
                 and kwargs.get("fused", True))
        steps = kwargs.get("max_steps", None)  # truncation hook used by benchmarks / tests; None = full chain
        noise_mode = kwargs.get("noise_mode", "batch")  # 'batch': randn_like(img) as the reference; 'shared': one draw per step
        idxs = list(range(self.num_timesteps))[::-1]
        if steps is not None:
            idxs = idxs[:steps]
        if fused:
            rec = dict(record_every=record_every, save_grids_path=kwargs.get("save_grids_path", None),
                       original_file_name=kwargs.get("original_file_name", "image_0")) if record else None
            return self._loop_fused(model, cond, x_start, measurement, sample_pattern, idxs, noise_mode, kwargs.get("progress"),
                                    kwargs.get("cuda_graph", True), record=rec)
        return self._loop_autograd(model, measurement_cond_fn, x_start, measurement, sample_pattern, idxs)

    def _draw(self, like, noise_mode):
        if noise_mode == "shared":
            return torch.randn((1,) + tuple(like.shape[1:]), device=like.device, dtype=like.dtype).expand_as(like).contiguous()
        return torch.randn_like(like)

    def _guidance_on(self, sample_pattern, idx):
        return (sample_pattern["pattern"] == "original") or (sample_pattern["pattern"] is None) or \
            (sample_pattern["start_guidance"] * self.num_timesteps >= idx >= sample_pattern["stop_guidance"] * self.num_timesteps)

    def fused_state(self, model, cond, x, measurement):
        """Device-resident buffers of the fused step (allocated once per loop)."""
        B, Cc, H, W = x.shape
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        if hasattr(cond, "check_batch"):
            cond.check_batch(x)
        return dict(
            coef=self.mean_processor.table.on(dev), t_idx=torch.zeros(B, dtype=torch.int32, device=dev),
            t_model=torch.zeros(B, **f32), freeze=torch.zeros(1, dtype=torch.int32, device=dev),
            model_out=torch.empty(B, 2 * Cc, H, W, **f32), x0=torch.empty(B, Cc, H, W, **f32),
            mean=torch.empty(B, Cc, H, W, **f32), logvar=torch.empty(B, Cc, H, W, **f32),
            g_x0=torch.empty(B, Cc, H, W, **f32), g_direct=torch.zeros(B, Cc, H, W, **f32),
            g_mo=torch.empty(B, 2 * Cc, H, W, **f32), g_unet=torch.zeros(B, Cc, H, W, **f32),
            zero_scale=torch.zeros(4, **f32), t_zero=torch.zeros(B, dtype=torch.int32, device=dev),
            ps_loss=torch.zeros(B, **f32),
            grad=torch.empty(B, Cc, H, W, **f32), losses=torch.zeros(B, 4, **f32),
            scale=cond._scale4(Cc).to(dev), y=measurement.contiguous().float().clone(),
            clip=(cond.gradient_clip_value if getattr(cond, "gradient_clip", False) else -1.0))

    def fused_step(self, model, cond, st, img, noise):
        """One guided reverse step on device buffers; `img` is updated in place.  No host synchronisation."""
        L = _lib.load()
        B, Cc, H, W = img.shape
        HW = H * W
        s = _lib.stream()
        flags = self.posterior_flags
        clip = bool(flags & 1)
        model._forward_raw(img, st["t_model"], out=st["model_out"])
        _lib.check(L.osm_posterior_fwd_ex(_lib.ptr(st["coef"]), _lib.ptr(st["t_idx"]), _lib.ptr(img), _lib.ptr(st["model_out"]),
                                          _lib.ptr(st["x0"]), _lib.ptr(st["mean"]), _lib.ptr(st["logvar"]), B, Cc, HW, flags, s))
        cond.guidance_gradient(st["x0"], st["y"], st["freeze"], st["g_x0"], st["losses"])
        _lib.check(L.osm_posterior_vjp_ex(_lib.ptr(st["coef"]), _lib.ptr(st["t_idx"]), _lib.ptr(st["g_x0"]), None, None,
                                          _lib.ptr(st["g_direct"]), _lib.ptr(st["g_mo"]), B, Cc, HW,
                                          _lib.ptr(img) if clip else None, _lib.ptr(st["model_out"]) if clip else None, flags, s))
        model._vjp_raw(st["g_mo"], grad_x=st["g_unet"])
        _lib.check(L.osm_sampler_update(_lib.ptr(st["mean"]), _lib.ptr(st["g_direct"]), _lib.ptr(st["g_unet"]),
                                        _lib.ptr(st["scale"]), st["clip"], _lib.ptr(st["logvar"]), _lib.ptr(noise),
                                        _lib.ptr(st["t_idx"]), _lib.ptr(img), _lib.ptr(st["grad"]), B, Cc, HW, s))

    def fused_step_unguided(self, model, st, img, noise):
        """A step with the guidance switched off (`guidance_flag` False, gaussian_diffusion.py:219-222, :262-264):
        img <- mean + exp(0.5 logvar) z.  Same kernels as the guided step minus guidance / VJP; the update kernel runs
        with a zero guidance scale (the gradient buffers are finite: zero-initialised or left by an earlier guided step)."""
        L = _lib.load()
        B, Cc, H, W = img.shape
        HW = H * W
        s = _lib.stream()
        model._forward_raw(img, st["t_model"], out=st["model_out"])
        _lib.check(L.osm_posterior_fwd_ex(_lib.ptr(st["coef"]), _lib.ptr(st["t_idx"]), _lib.ptr(img), _lib.ptr(st["model_out"]),
                                          _lib.ptr(st["x0"]), _lib.ptr(st["mean"]), _lib.ptr(st["logvar"]), B, Cc, HW,
                                          self.posterior_flags, s))
        _lib.check(L.osm_sampler_update(_lib.ptr(st["mean"]), _lib.ptr(st["g_direct"]), _lib.ptr(st["g_unet"]),
                                        _lib.ptr(st["zero_scale"]), -1.0, _lib.ptr(st["logvar"]), _lib.ptr(noise),
                                        _lib.ptr(st["t_idx"]), _lib.ptr(img), None, B, Cc, HW, s))

    # DDIM's eta (gaussian_diffusion.py:507); None = this sampler is ancestral (DDPM.p_sample, :494-503)
    ddim_eta = None

    def fused_step_ps(self, model, cond, st, img, noise):
        """One step of the rgb_guidance branch on device buffers: p_sample (DDPM or DDIM, with clip_denoised when the
        sampler asks for it), the `ps` norm gradient, the UNet input-VJP, and x_t <- sample - scale * grad."""
        L = _lib.load()
        B, Cc, H, W = img.shape
        HW = H * W
        s = _lib.stream()
        flags = self.posterior_flags
        clip = flags & 1
        model._forward_raw(img, st["t_model"], out=st["model_out"])
        _lib.check(L.osm_posterior_fwd_ex(_lib.ptr(st["coef"]), _lib.ptr(st["t_idx"]), _lib.ptr(img), _lib.ptr(st["model_out"]),
                                          _lib.ptr(st["x0"]), _lib.ptr(st["mean"]), _lib.ptr(st["logvar"]), B, Cc, HW, flags, s))
        base, t_noise = st["mean"], st["t_idx"]
        if self.ddim_eta is not None:   # the DDIM sample replaces mean + sigma z; the update kernel then adds no noise
            _lib.check(L.osm_ddim_sample(_lib.ptr(st["coef"]), _lib.ptr(st["t_idx"]), _lib.ptr(img), _lib.ptr(st["x0"]),
                                         _lib.ptr(noise), float(self.ddim_eta), _lib.ptr(st["mean"]), B, Cc, HW, s))
            t_noise = st["t_zero"]
        cond.guidance_gradient(st["x0"], st["y"], st["g_x0"], st["ps_loss"])
        _lib.check(L.osm_posterior_vjp_ex(_lib.ptr(st["coef"]), _lib.ptr(st["t_idx"]), _lib.ptr(st["g_x0"]), None, None,
                                          _lib.ptr(st["g_direct"]), _lib.ptr(st["g_mo"]), B, Cc, HW,
                                          _lib.ptr(img) if clip else None, _lib.ptr(st["model_out"]) if clip else None, flags, s))
        model._vjp_raw(st["g_mo"], grad_x=st["g_unet"])
        _lib.check(L.osm_sampler_update_ex(_lib.ptr(base), _lib.ptr(st["g_direct"]), _lib.ptr(st["g_unet"]),
                                           _lib.ptr(st["scale"]), -1.0, _lib.ptr(st["logvar"]), _lib.ptr(noise),
                                           _lib.ptr(t_noise), _lib.ptr(img), _lib.ptr(st["grad"]), B, Cc, HW, 1, s))

    def _loop_fused_ps(self, model, cond, x_start, measurement, idxs, noise_mode, progress=None, cuda_graph=True):
        host_meas = None
        if not measurement.is_cuda:
            host_meas = measurement.contiguous().float()
            if not host_meas.is_pinned():
                host_meas = host_meas.pin_memory()
        stepper = self._stepper_for(model, cond, x_start, measurement if host_meas is None else host_meas, None, noise_mode,
                                    cuda_graph)
        st = stepper.st
        for idx in idxs:
            if host_meas is not None:
                st["y"].copy_(host_meas, non_blocking=True)
            stepper.step(idx)
            if progress is not None:
                progress(idx, st["ps_loss"].cpu().numpy())
        self.last_loss = st["ps_loss"].clone()
        self.last_gradients = st["grad"].clone()
        self.last_pred_xstart = st["x0"].clone()
        return stepper.img.clone()

    def _loop_autograd_ps(self, model, measurement_cond_fn, x_start, measurement, idxs):
        img = x_start
        for idx in idxs:
            time = torch.tensor([idx] * img.shape[0], device=img.device)
            img = img.requires_grad_()
            out = self.p_sample(x=img, t=time, model=model)
            noisy_measurement = self.q_sample(measurement, t=time)     # dead draw, after p_sample's (:241)
            img, loss = measurement_cond_fn(x_t=out["sample"], measurement=measurement, noisy_measurement=noisy_measurement,
                                            x_prev=img, x_0_hat=out["pred_xstart"])
            img = img.detach_()
            self.last_loss = loss
        return img

    def _loop_fused(self, model, cond, x_start, measurement, sample_pattern, idxs, noise_mode, progress=None, cuda_graph=True,
                    record=None):
        """`measurement` may live in (pinned) host memory: it is then streamed to the device every step.  `progress(idx,
        loss[B] numpy)` is called after every step when given - it costs one device->host read per step, which is what
        the reference's progress bar does (gaussian_diffusion.py:276-296)."""
        host_meas = None
        if not measurement.is_cuda:
            host_meas = measurement.contiguous().float()
            if not host_meas.is_pinned():
                host_meas = host_meas.pin_memory()
        stepper = self._stepper_for(model, cond, x_start, measurement if host_meas is None else host_meas, sample_pattern,
                                    noise_mode, cuda_graph)
        st, img = stepper.st, stepper.img
        # progress read-back is deferred by one step: the loss of step k is copied to a pinned slot on the stream and
        # handed to the callback after step k+1 has been enqueued, so the device never waits for the host.
        pend, slots, slot_ev = None, None, None
        if progress is not None:
            slots = [torch.empty(st["losses"].shape[0], dtype=torch.float32).pin_memory() for _ in range(2)]
            slot_ev = [torch.cuda.Event() for _ in range(2)]
        for k, idx in enumerate(idxs):
            if host_meas is not None:
                st["y"].copy_(host_meas, non_blocking=True)
            # alternate length M of the gibbsDDRM-style pattern (gaussian_diffusion.py:224-227): M full steps at this index
            for _ in range(utilso.set_alternate_length(sample_pattern, idx, self.num_timesteps)):
                stepper.step(idx)
            if progress is not None:
                slots[k & 1].copy_(st["losses"][:, 0], non_blocking=True)
                slot_ev[k & 1].record()
                if pend is not None:
                    slot_ev[pend[0]].synchronize()
                    progress(pend[1], slots[pend[0]].numpy().copy())
                pend = (k & 1, idx)
            # `record` (gaussian_diffusion.py:310-326): keep pred_xstart of image 0 every record_every steps.  The snapshot
            # is a device-to-device copy on the stream (no host sync in the loop); images are built after the loop.
            if record is not None and ((idx % record["record_every"] == 0) or (idx == 0) or (idx == 999)):
                record.setdefault("frames", []).append(st["x0"][0:1].clone())
        if pend is not None:
            slot_ev[pend[0]].synchronize()
            progress(pend[1], slots[pend[0]].numpy().copy())
        if record is not None:
            self.last_record = self._finish_record(record)
        variable_dict = cond.operator.optimize(freeze_phi=True)
        loss = st["losses"][:, 0].cpu().numpy()
        self.last_gradients = st["grad"].clone()    # the stepper (and its buffers) is reused by the next p_sample_loop call
        self.last_aux = st["losses"].clone()
        return img.clone(), variable_dict, loss, st["x0"].detach().cpu()

    def _stepper_for(self, model, cond, x_start, measurement, sample_pattern, noise_mode, cuda_graph):
        """One FusedStepper (device state + captured CUDA graph) per (model, conditioner, batch shape): consecutive
        `p_sample_loop` calls - the reference's per-image loop, `osmosis_sampling.py:117-199` - reuse the captured graph and
        only refresh the image, the measurement and the step scalars."""
        cache = self.__dict__.setdefault("_steppers", {})
        key = (id(model), id(cond), tuple(x_start.shape), str(x_start.device), noise_mode, bool(cuda_graph),
               tuple(sorted((k, str(v)) for k, v in (sample_pattern or {}).items())))
        stepper = cache.get(key)
        if stepper is None:
            img = x_start.detach().clone().contiguous().float()
            y = torch.empty(measurement.shape, dtype=torch.float32, device=img.device)
            y.copy_(measurement, non_blocking=True)
            stepper = FusedStepper(self, model, cond, img, y, sample_pattern, noise_mode=noise_mode, cuda_graph=cuda_graph)
            while len(cache) >= 2:
                cache.pop(next(iter(cache)))
            cache[key] = stepper
        else:
            stepper.img.copy_(x_start.detach())
            stepper.st["y"].copy_(measurement, non_blocking=True)
        return stepper

    @staticmethod
    def _finish_record(record):
        """Builds the reference's process grid (gaussian_diffusion.py:310-333) from the recorded pred_xstart frames: clipped
        RGB row over a percentile-normalised (0.05 / 0.99), colour-mapped depth row; saved as <name>_process.png when a
        grid path is given.  Returns dict(rgb=[n,3,H,W], depth_color=[n,3,H,W]) on the device (+ 'grid' on the host)."""
        frames = torch.cat(record.get("frames", []), 0) if record.get("frames") else None
        if frames is None:
            return None
        rgb = torch.clamp(0.5 * (frames[:, 0:3] + 1), 0, 1)
        depth_pmm = utilso.min_max_norm_range_percentile(frames[:, 3:4].contiguous(), percent_low=0.05, percent_high=0.99)
        depth_color = utilso.depth_tensor_to_color_image(depth_pmm) if frames.shape[0] > 1 else \
            utilso.depth_tensor_to_color_image(depth_pmm)[None]
        out = dict(rgb=rgb, depth_color=depth_color)
        if record.get("save_grids_path") is not None:
            from torchvision.utils import make_grid
            import torchvision.transforms.functional as tvtf
            import os
            grid = make_grid(list(rgb.cpu()) + list(depth_color.cpu()), nrow=rgb.shape[0])
            tvtf.to_pil_image(grid).save(os.path.join(record["save_grids_path"], f"{record['original_file_name']}_process.png"))
            out["grid"] = grid
        return out

    def _loop_autograd(self, model, measurement_cond_fn, x_start, measurement, sample_pattern, idxs):
        img = x_start
        device = img.device
        loss, variable_dict, out = None, {}, None
        for idx in idxs:
            time = torch.tensor([idx] * img.shape[0], device=device)
            guidance = self._guidance_on(sample_pattern, idx)
            for _ in range(utilso.set_alternate_length(sample_pattern, idx, self.num_timesteps)):
                img.requires_grad_(bool(guidance))
                out = self.p_mean_variance(model=model, x=img, t=time)
                out["sample"] = out["mean"]
                noisy_measurement = self.q_sample(measurement, t=time)
                freeze = utilso.is_freeze_phi(sample_pattern, idx, self.num_timesteps)
                if guidance:
                    img, loss, variable_dict, _grads, _aux = measurement_cond_fn(
                        x_t=out["sample"], measurement=measurement, noisy_measurement=noisy_measurement, x_prev=img,
                        x_0_hat=out["pred_xstart"], freeze_phi=freeze, time_index=float(idx) / self.num_timesteps)
                else:
                    img = out["sample"]
                noise = torch.randn_like(img)
                img = img.detach()
                if idx != 0:
                    img = img + torch.exp(0.5 * out["log_variance"].detach()) * noise
        return img, variable_dict, loss, out["pred_xstart"].detach().cpu()


class FusedStepper:
    """Runs guided reverse steps on device-resident state, replaying ONE captured CUDA graph per step.

    Everything a step needs that changes from step to step (respaced index, model timestep, freeze flag, noise, phi,
    the image itself) lives in fixed device buffers, so the ~850-kernel launch sequence of a step is captured once
    (after one eager warm-up step) and replayed; per step the host only refreshes three scalars and draws the noise
    with torch's generator (same draws, same order as the reference: the dead q_sample draw, then the step noise)."""

    def __init__(self, sampler, model, cond, img, measurement, sample_pattern, noise_mode="batch", cuda_graph=True):
        self.sampler, self.model, self.cond, self.img = sampler, model, cond, img
        self.sample_pattern, self.noise_mode = sample_pattern, noise_mode
        self.st = sampler.fused_state(model, cond, img, measurement)
        self.noise = torch.empty_like(img)
        self.dead = torch.empty_like(self.st["y"])
        from .condition_methods import PosteriorSampling
        self.ps = isinstance(cond, PosteriorSampling)   # rgb_guidance branch: p_sample + `ps` conditioning
        self.use_graph = bool(cuda_graph)
        self.graph = None
        self.graph_unguided = None
        self.calls = 0
        self.calls_unguided = 0
        self.bind_gen = None

    def _check_plan(self):
        """The captured graphs hold raw pointers into the UNet engine's bound workspace.  `UNetModel._ensure_bound` frees and
        re-plans it whenever another (B, H, W) runs through the same model, so graphs captured under an older plan are
        dropped here (the next step runs eagerly, which re-binds, and the one after re-captures)."""
        m = self.model
        if self.bind_gen is not None and (m._bind_generation != self.bind_gen or m._bound != (self.img.shape[0],) + tuple(self.img.shape[2:])):
            self.graph = self.graph_unguided = None
            self.calls = self.calls_unguided = 0
        self.bind_gen = None

    def _draw_into(self, buf):
        if self.noise_mode == "shared":
            buf.copy_(torch.randn((1,) + tuple(buf.shape[1:]), device=buf.device, dtype=buf.dtype).expand_as(buf))
        else:
            buf.normal_()

    def step(self, idx, freeze=None):
        s, st, T = self.sampler, self.st, self.sampler.num_timesteps
        self._check_plan()
        try:
            self._step(idx, freeze)
        finally:
            self.bind_gen = self.model._bind_generation

    def _step(self, idx, freeze=None):
        s, st, T = self.sampler, self.st, self.sampler.num_timesteps
        if self.ps:
            st["t_idx"].fill_(idx)
            st["t_model"].fill_(s._model_timestep(idx))
            self._draw_into(self.noise)    # p_sample's draw comes first in this branch (gaussian_diffusion.py:498 / :522),
            self._draw_into(self.dead)     # then the dead q_sample draw (:241)
            if self.use_graph and self.calls >= 1:
                if self.graph is None:
                    torch.cuda.synchronize()
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph):
                        s.fused_step_ps(self.model, self.cond, st, self.img, self.noise)
                self.graph.replay()
            else:
                s.fused_step_ps(self.model, self.cond, st, self.img, self.noise)
            self.calls += 1
            return
        guided = s._guidance_on(self.sample_pattern, idx)
        if freeze is None:
            freeze = utilso.is_freeze_phi(self.sample_pattern, idx, T)
        st["t_idx"].fill_(idx)
        st["t_model"].fill_(s._model_timestep(idx))
        st["freeze"].fill_(1 if freeze else 0)
        self.cond.operator.set_variable_gradients(value=not freeze)
        self._draw_into(self.dead)     # dead q_sample draw: RNG parity with gaussian_diffusion.py:241
        self._draw_into(self.noise)    # drawn even at t = 0 (:266)
        if not guided:
            if self.use_graph and self.calls_unguided >= 1:
                if self.graph_unguided is None:
                    torch.cuda.synchronize()
                    self.graph_unguided = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph_unguided):
                        s.fused_step_unguided(self.model, st, self.img, self.noise)
                self.graph_unguided.replay()
            else:
                s.fused_step_unguided(self.model, st, self.img, self.noise)
            self.calls_unguided += 1
            return
        if self.use_graph and self.calls >= 1:
            if self.graph is None:
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    s.fused_step(self.model, self.cond, st, self.img, self.noise)
            self.graph.replay()
        else:
            s.fused_step(self.model, self.cond, st, self.img, self.noise)
        self.calls += 1


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process that keeps a subset of the base process' timesteps (betas re-derived from abar)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base_abar = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64), axis=0)
        last, new_betas = 1.0, []
        for i, abar in enumerate(base_abar):
            if i in self.use_timesteps:
                new_betas.append(1 - abar / last)
                last = abar
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t  # done by the wrapped model

    def _model_timestep(self, idx):
        t = float(self.timestep_map[idx])
        return t * (1000.0 / self.original_num_steps) if self.rescale_timesteps else t


class _WrappedModel:
    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._map = {}

    def __call__(self, x, ts, **kwargs):
        key = (str(ts.device), ts.dtype)
        if key not in self._map:
            self._map[key] = torch.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        new_ts = self._map[key][ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)


@register_sampler(name="ddpm")
class DDPM(SpacedDiffusion):
    def p_sample(self, model, x, t):
        out = self.p_mean_variance(model, x, t)
        sample = out["mean"]
        noise = torch.randn_like(x)
        if t[0] != 0:
            sample = sample + torch.exp(0.5 * out["log_variance"]) * noise
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}


@register_sampler(name="ddim")
class DDIM(SpacedDiffusion):
    """gaussian_diffusion.py:505-535.  In the fused loop the sample is one kernel (osm_ddim_sample); `p_sample` below is the
    reference-facing, autograd-compatible form of the same arithmetic."""
    ddim_eta = 0.0

    def p_sample(self, model, x, t, eta=0.0):
        out = self.p_mean_variance(model, x, t)
        L = _lib.load()
        B, Cc, H, W = x.shape
        noise = torch.randn_like(x)
        sample = torch.empty_like(out["pred_xstart"])
        coef = self.mean_processor.table.on(x.device)
        _lib.check(L.osm_ddim_sample(_lib.ptr(coef), _lib.ptr(t.to(torch.int32).contiguous()), _lib.ptr(x.detach().contiguous()),
                                     _lib.ptr(out["pred_xstart"].detach().contiguous()), _lib.ptr(noise), float(eta),
                                     _lib.ptr(sample), B, Cc, H * W, _lib.stream()))
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    def predict_eps_from_x_start(self, x_t, t, pred_xstart):
        coef1 = extract_and_expand(self.sqrt_recip_alphas_cumprod, t, x_t)
        coef2 = extract_and_expand(self.sqrt_recipm1_alphas_cumprod, t, x_t)
        return (coef1 * x_t - pred_xstart) / coef2
