"""Conditioning (guidance) methods.

Mirrors the reference `guided_diffusion/condition_methods.py`: the registry (:8-24), `ConditioningMethod`
(:27-58) and `PosteriorSamplingOsmosis` / name `osmosis` (:61-231) with the options the shipped configs
select: `loss_function: norm | mse`, `loss_weight: depth | none`, `gradient_x_prev: True`, gradient clipping, the
auxiliary losses and the pcgs sampling pattern; and `PosteriorSampling` / name `ps` (:234-251), plain DPS with the
gaussian branch of `ConditioningMethod.grad_and_value` (:36-40) - what rgb_guidance_sample_config.yaml selects.
`gradient_x_prev: False` raises (it does inside the reference too); the poisson branch (:42-48) is not built.

What the reference does in ~66 small ATen kernels and 3 device->host syncs per inner iteration
(grad_and_value :109-144, AuxiliaryLoss, backward, operator.optimize) is ONE launch here
(osm_guidance_phi_loop): all `n_iter` evaluations, the phi SGD steps and d(loss)/d(x_0_hat).  The gradient
then flows to `x_prev` through whatever produced `x_0_hat` - the native UNet's input-VJP when the model is
ours (autograd Function), or plain autograd for a foreign model.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import lib as _lib
from ..osmosis_utils import losses as losseso
from ..osmosis_utils import utils as utilso
from .measurements import DEPTH_KIND, OP_KIND, depth_spec

__CONDITIONING_METHOD__ = {}


def register_conditioning_method(name: str):
    def wrapper(cls):
        if __CONDITIONING_METHOD__.get(name, None):
            raise NameError(f"Name {name} is already registered!")
        __CONDITIONING_METHOD__[name] = cls
        return cls
    return wrapper


def get_conditioning_method(name: str, operator, noiser, **kwargs):
    if __CONDITIONING_METHOD__.get(name, None) is None:
        raise NameError(f"Name {name} is not defined!")
    return __CONDITIONING_METHOD__[name](operator=operator, noiser=noiser, **kwargs)


class ConditioningMethod:
    def __init__(self, operator, noiser, **kwargs):
        self.operator = operator
        self.noiser = noiser

    def conditioning(self, x_prev, x_t, x_0_hat, measurement, **kwargs):
        raise NotImplementedError


def _parse_scale(s):
    try:
        return [float(s)]
    except ValueError:
        return [float(v.strip()) for v in s.split(",")]


@register_conditioning_method(name="osmosis")
class PosteriorSamplingOsmosis(ConditioningMethod):
    def __init__(self, operator, noiser, **kwargs):
        super().__init__(operator, noiser)
        self.scale = torch.tensor(_parse_scale(kwargs.get("scale", 1.0)))
        self.gradient_x_prev = kwargs.get("gradient_x_prev", False)
        self.pattern_name = kwargs.get("pattern", "original")
        self.global_N = kwargs.get("global_N", 1)
        self.local_M = kwargs.get("local_M", 1)
        self.n_iter = kwargs.get("n_iter", 1)
        self.update_start = kwargs.get("update_start", 1.0)
        aux = kwargs.get("aux_loss", None)
        self.aux_loss = losseso.AuxiliaryLoss({k: float(v) for k, v in aux.items()}) if aux is not None else None
        self.loss_function = kwargs.get("loss_function", "norm")
        self.loss_weight = kwargs.get("loss_weight", None)
        self.weight_function = kwargs.get("weight_function", None)
        clip = [s for s in kwargs.get("gradient_clip", "False").split(",")]
        self.gradient_clip = utilso.str2bool(clip[0])
        self.gradient_clip_value = float(clip[1].strip()) if self.gradient_clip else None
        if self.loss_function not in ("norm", "mse"):
            raise NotImplementedError                       # as the reference does at call time (:140-142)
        if not self.gradient_x_prev:
            # the reference's branch cannot run: it detaches x_0_hat, switches x_prev.requires_grad off and then calls
            # total_loss.backward(inputs=[x_prev]) (:152-156, :186-191), which torch rejects
            raise NotImplementedError("gradient_x_prev: False raises inside the reference as well (backward w.r.t. a tensor "
                                      "that does not require grad)")
        self._params = None
        self._dev = {}

    # ---- kernel parameter block ---------------------------------------------------------------------
    def kernel_params(self):
        if self._params is None:
            op = self.operator
            kind = getattr(op, "kind_name", None)
            if kind not in OP_KIND:
                raise NotImplementedError(f"the native `osmosis` conditioning has guidance kernels for the operators {sorted(OP_KIND)}; "
                                          f"{type(op).__name__} (kind {kind!r}) is not one of them")
            p = _lib.GuidanceParamsC()
            p.op_kind = OP_KIND[kind]
            p.depth_kind = op.depth_kind
            for i in range(3):
                p.depth_val[i] = op.depth_val[i]
                p.eta[i] = op.eta[i] if i < len(op.eta) else 0.0
            if self.loss_weight in (None, "none"):
                p.weight_kind, p.weight_depth_kind = 0, 0
            elif self.loss_weight == "depth":
                parts = self.weight_function.split(",") if isinstance(self.weight_function, str) else ["none"]
                fn = parts[0]
                if fn not in DEPTH_KIND:
                    raise NotImplementedError
                kind, vals = depth_spec(fn, ",".join(parts[1:]) if len(parts) > 1 else None)
                p.weight_kind, p.weight_depth_kind = 1, kind
                for i in range(3):
                    p.weight_val[i] = vals[i]
            else:
                raise NotImplementedError
            p.n_iter = int(self.n_iter)
            w = self.aux_loss.kernel_weights() if self.aux_loss is not None else {"gamma_avrg": 0.0, "gamma_val": 0.0}
            p.gamma_avrg, p.gamma_val = w["gamma_avrg"], w["gamma_val"]
            p.loss_kind = 1 if self.loss_function == "mse" else 0
            p.optimizer = 1 if str(op.optimizer).lower() == "adam" else 0
            p.opt_state = op.opt_state.data_ptr() if op.opt_state is not None else None
            p.phi_batch = int(op.phi.shape[0])
            self._params = p
        return self._params

    def _buffers(self, x):
        key = (str(x.device), tuple(x.shape))
        if key not in self._dev:
            B = x.shape[0]
            self._dev = {key: dict(freeze=torch.zeros(1, dtype=torch.int32, device=x.device),
                                   g_x0=torch.empty_like(x), losses=torch.zeros(B, 4, dtype=torch.float32, device=x.device),
                                   zero_t=torch.zeros(B, dtype=torch.int32, device=x.device),
                                   scale=self._scale4(x.shape[1]).to(x.device))}
        return self._dev[key]

    def _scale4(self, channels):
        s = self.scale.float()
        return (s.repeat(channels) if s.numel() == 1 else s).contiguous()

    def check_batch(self, x):
        """The operator's phi / optimizer state hold one row per image (`batch_size` of get_operator, which the reference takes
        from data.batch_size = 1 in every shipped YAML): a batch of another size would index them out of bounds."""
        n = int(self.operator.phi.shape[0])
        if int(x.shape[0]) != n:
            raise ValueError(f"the operator was built for batch_size={n} but the sampler runs {int(x.shape[0])} images: pass "
                             f"batch_size={int(x.shape[0])} to get_operator")

    def guidance_gradient(self, x_0_hat, measurement, freeze_flag_dev, g_x0, losses):
        """One launch: phi loop + d(total loss)/d(x_0_hat).  All arguments are device tensors."""
        self.check_batch(x_0_hat)
        B, Cc, H, W = x_0_hat.shape
        L = _lib.load()
        _lib.check(L.osm_guidance_phi_loop(C.byref(self.kernel_params()), _lib.ptr(x_0_hat), _lib.ptr(measurement),
                                           _lib.ptr(self.operator.phi), _lib.ptr(freeze_flag_dev), _lib.ptr(g_x0),
                                           _lib.ptr(losses), B, H * W, _lib.stream()))

    # ---- the reference-facing call (autograd-compatible path) ------------------------------------------
    def conditioning(self, x_prev, x_t, x_0_hat, measurement, **kwargs):
        freeze_phi = kwargs.get("freeze_phi", False)
        self.check_batch(x_t)
        buf = self._buffers(x_t)
        self.operator.set_variable_gradients(value=not freeze_phi)
        buf["freeze"].fill_(1 if freeze_phi else 0)
        x0 = x_0_hat.detach().contiguous()
        self.guidance_gradient(x0, measurement.contiguous(), buf["freeze"], buf["g_x0"], buf["losses"])
        variables_dict = self.operator.optimize(freeze_phi=freeze_phi)
        # d loss / d x_prev through the graph that produced x_0_hat (the UNet's input-VJP)
        if x_prev.grad is not None:
            x_prev.grad = None
        torch.autograd.backward([x_0_hat], [buf["g_x0"]], inputs=[x_prev])
        grad = x_prev.grad
        with torch.no_grad():
            L = _lib.load()
            B, Cc, H, W = x_t.shape
            clip = self.gradient_clip_value if self.gradient_clip else -1.0
            xt = x_t.detach()
            assert xt.is_contiguous()
            # x_t -= scale * clamp(grad): the sampler-update kernel with the noise term switched off (t_idx = 0)
            _lib.check(L.osm_sampler_update(_lib.ptr(xt), _lib.ptr(grad.contiguous()), None, _lib.ptr(buf["scale"]), clip,
                                            _lib.ptr(xt), _lib.ptr(xt), _lib.ptr(buf["zero_t"]), _lib.ptr(xt), None, B, Cc,
                                            H * W, _lib.stream()))
        losses = buf["losses"].cpu()
        sep_loss = losses[:, 0].numpy()
        aux_loss_dict = None
        if self.aux_loss is not None:
            cols = {"avrg_loss": 1, "val_loss": 2}
            aux_loss_dict = {k: losses[:, cols[k]].clone() for k in self.aux_loss.losses_dictionary}
        return x_t, sep_loss, variables_dict, grad, aux_loss_dict


@register_conditioning_method(name="ps")
class PosteriorSampling(ConditioningMethod):
    """condition_methods.py:234-251: x_t <- x_t - scale_c * d||y - A(x_0_hat[:, :3])||_2 / d x_prev with a parameter-free
    operator.  Native for the identity operators (`rgb_guidance`, `noise`): the norm and its gradient w.r.t. x_0_hat are ONE
    launch (osm_ps_guidance); the gradient then flows to x_prev through whatever produced x_0_hat (the native UNet's
    input-VJP).  Per-image norm: a batch of B is B independent reference runs (the reference takes one norm over the
    whole batch, which couples the images; it is only ever run at B = 1)."""

    def __init__(self, operator, noiser, **kwargs):
        super().__init__(operator, noiser)
        self.scale = torch.tensor(_parse_scale(kwargs.get("scale", 1.0)))
        if getattr(noiser, "__name__", "gaussian") not in ("gaussian", "clean"):
            raise NotImplementedError("the `ps` conditioning is built for the gaussian branch (condition_methods.py:36-40)")
        if not getattr(operator, "is_identity", False) and not type(operator).__name__ == "DenoiseOperator":
            raise NotImplementedError("the native `ps` conditioning supports the identity operators (rgb_guidance, noise)")
        self._dev = {}

    _scale4 = PosteriorSamplingOsmosis._scale4

    def guidance_gradient(self, x_0_hat, measurement, g_x0, losses):
        B, Cc, H, W = x_0_hat.shape
        _lib.check(_lib.load().osm_ps_guidance(_lib.ptr(x_0_hat), _lib.ptr(measurement), _lib.ptr(g_x0), _lib.ptr(losses), B, Cc,
                                               H * W, _lib.stream()))

    def grad_and_value(self, x_prev, x_0_hat, measurement, **kwargs):
        x0 = x_0_hat.detach().contiguous()
        g_x0 = torch.empty_like(x0)
        losses = torch.empty(x0.shape[0], dtype=torch.float32, device=x0.device)
        self.guidance_gradient(x0, measurement.contiguous().float(), g_x0, losses)
        if x_prev.grad is not None:
            x_prev.grad = None
        torch.autograd.backward([x_0_hat], [g_x0], inputs=[x_prev])
        return x_prev.grad, losses

    def conditioning(self, x_prev, x_t, x_0_hat, measurement, **kwargs):
        norm_grad, norm = self.grad_and_value(x_prev=x_prev, x_0_hat=x_0_hat, measurement=measurement, **kwargs)
        B, Cc, H, W = x_t.shape
        key = (str(x_t.device), B, Cc)
        if key not in self._dev:
            self._dev = {key: dict(zero_t=torch.zeros(B, dtype=torch.int32, device=x_t.device), scale=self._scale4(Cc).to(x_t.device))}
        buf = self._dev[key]
        with torch.no_grad():
            xt = x_t.detach().contiguous()
            # x_t -= grad * scale: the update kernel with clamping and the noise term switched off
            _lib.check(_lib.load().osm_sampler_update_ex(_lib.ptr(xt), _lib.ptr(norm_grad.contiguous()), None, _lib.ptr(buf["scale"]),
                                                         -1.0, _lib.ptr(xt), _lib.ptr(xt), _lib.ptr(buf["zero_t"]), _lib.ptr(xt), None,
                                                         B, Cc, H * W, 1, _lib.stream()))
        return xt, (norm[0] if B == 1 else norm)
