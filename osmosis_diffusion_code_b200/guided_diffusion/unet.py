"""`create_model` / `UNetModel` backed by the native sm_100a engine.

Mirrors the reference factory `guided_diffusion/unet.py:create_model` (:27-98): same keyword arguments
(unknown ones are rejected exactly like a Python signature would), same channel_mult / attention_resolutions
parsing, the osmosis 4-in / 8-out surgery (`osmosis_utils/utils.py:265-288`), and the "try to load the
checkpoint, print and keep the random init on failure" behaviour (:94-97).  The returned object is called as
`model(x[B,4,H,W], t[B]) -> [B,8,H,W]` (unet.py:713-742) and records an autograd edge whose backward is the
engine's input-VJP, so `total_loss.backward(inputs=[x_prev] + phis)` (condition_methods.py:186-191) works
on it unchanged.  Weight gradients do not exist: the sampling path never asks for them.

Only the configuration family the shipped YAMLs use is supported natively (scale-shift norm, ResBlock
up/down, legacy attention order, fp32 storage, no class conditioning); anything else raises.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from .. import lib as _lib

NUM_CLASSES = 1000


def _parse_channel_mult(channel_mult, image_size):
    if channel_mult == "":
        table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}
        if image_size not in table:
            raise ValueError(f"unsupported image size: {image_size}")
        return table[image_size]
    if isinstance(channel_mult, str):
        return tuple(int(v) for v in channel_mult.split(","))
    return tuple(channel_mult)


def create_model(image_size, num_channels, num_res_blocks, channel_mult="", learn_sigma=False, class_cond=False,
                 use_checkpoint=False, attention_resolutions="16", num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0, resblock_updown=False, use_fp16=False,
                 use_new_attention_order=False, model_path="", pretrain_model="", conv_mode="tc"):
    mult = _parse_channel_mult(channel_mult, image_size)
    if isinstance(attention_resolutions, int):
        ds = (image_size // attention_resolutions,)
    elif isinstance(attention_resolutions, str):
        ds = tuple(image_size // int(r) for r in attention_resolutions.split(","))
    else:
        raise NotImplementedError
    in_ch, out_ch = 3, (6 if learn_sigma else 3)
    if pretrain_model == "osmosis":  # change_input_output_unet(model, 4, 8)
        in_ch, out_ch = 4, 8
    model = UNetModel(image_size=image_size, in_channels=in_ch, model_channels=num_channels, out_channels=out_ch,
                      num_res_blocks=num_res_blocks, attention_resolutions=ds, dropout=dropout, channel_mult=mult,
                      num_classes=(NUM_CLASSES if class_cond else None), use_checkpoint=use_checkpoint, use_fp16=use_fp16,
                      num_heads=num_heads, num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                      use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                      use_new_attention_order=use_new_attention_order, conv_mode=conv_mode)
    model._surgery = pretrain_model == "osmosis"
    try:
        model.load_state_dict(torch.load(model_path, map_location="cpu"))
    except Exception as e:  # same policy as the reference: report and continue with the random init
        print(f"Got exception: {e} / Randomly initialize")
    return model


class _UNetFn(torch.autograd.Function):
    """Autograd edge around the engine: forward = osm_unet_forward, backward = osm_unet_vjp_input."""

    @staticmethod
    def forward(ctx, x, t, model):
        out = model._forward_raw(x, t)
        ctx.model = model
        ctx.generation = model._generation
        return out

    @staticmethod
    def backward(ctx, grad_out):
        m = ctx.model
        if ctx.generation != m._generation:
            raise RuntimeError("the UNet engine keeps activations of the most recent forward only; "
                               "backward through an older call is not possible")
        return m._vjp_raw(grad_out.contiguous()), None, None


class UNetModel:
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, conv_mode="tc"):
        unsupported = []
        if not use_scale_shift_norm: unsupported.append("use_scale_shift_norm=False")
        if not resblock_updown: unsupported.append("resblock_updown=False")
        if use_new_attention_order: unsupported.append("use_new_attention_order=True")
        if use_fp16: unsupported.append("use_fp16=True")
        if num_classes is not None: unsupported.append("class_cond=True")
        if dims != 2: unsupported.append(f"dims={dims}")
        if dropout: unsupported.append("dropout>0 (sampling runs in eval mode)")
        if any(int(m) != m for m in channel_mult): unsupported.append("fractional channel_mult")
        if unsupported:
            raise NotImplementedError("native UNet engine: unsupported options: " + ", ".join(unsupported))
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions, self.channel_mult = tuple(attention_resolutions), tuple(int(m) for m in channel_mult)
        self.num_heads, self.num_head_channels = num_heads, num_head_channels
        self.dtype = torch.float32
        self.conv_mode = {"tc": 0, "tf32": 0, "fp32": 1, "exact": 1}[conv_mode] if isinstance(conv_mode, str) else int(conv_mode)
        L = _lib.load()
        cfg = _lib.UNetConfigC()
        cfg.in_channels, cfg.out_channels, cfg.model_channels = in_channels, out_channels, model_channels
        cfg.num_res_blocks, cfg.num_levels = num_res_blocks, len(self.channel_mult)
        for i, m in enumerate(self.channel_mult): cfg.channel_mult[i] = m
        cfg.num_attention_ds = len(self.attention_resolutions)
        for i, d in enumerate(self.attention_resolutions): cfg.attention_ds[i] = int(d)
        cfg.num_heads, cfg.num_head_channels, cfg.conv_mode = num_heads, num_head_channels, self.conv_mode
        h = C.c_void_p()
        _lib.check(L.osm_unet_create(C.byref(cfg), C.byref(h)))
        self._h, self._L = h, L
        self._specs = []
        name, nd, shp = C.c_char_p(), C.c_int(), (C.c_int64 * 4)()
        for i in range(L.osm_unet_param_count(h)):
            _lib.check(L.osm_unet_param_info(h, i, C.byref(name), C.byref(nd), shp))
            self._specs.append((name.value.decode(), tuple(int(shp[k]) for k in range(nd.value))))
        self._host = None        # {name: CPU fp32 tensor}
        self.device = None
        self._uploaded = False
        self._bound = None       # (B, H, W)
        self._ws = None
        self._generation = 0
        self._surgery = False    # set by create_model for pretrain_model == "osmosis" (change_input_output_unet)
        self._bind_generation = 0   # bumped whenever the workspace is re-planned: captured CUDA graphs of older plans are stale

    # ---- parameters -------------------------------------------------------------------------------
    def param_specs(self):
        return list(self._specs)

    def num_params(self):
        return sum(math.prod(s) for _, s in self._specs)

    def _default_init(self):
        """Default-initialised state_dict (what the reference is left with when the checkpoint is missing):
        uniform(+-1/sqrt(fan_in)) convs / linears, unit GroupNorm, and zeros where it applies zero_module."""
        g = torch.Generator().manual_seed(0)
        sd = {}
        shapes = dict(self._specs)
        for name, shape in self._specs:
            stem, leaf = name.rsplit(".", 1)
            # the osmosis 4-in / 8-out surgery (utils.py:279, :286) replaces out[-1] by a FRESH nn.Conv2d: default init, not zero
            zeroed = ".out_layers.3." in name or ".proj_out." in name or (name.startswith("out.2.") and not self._surgery)
            if ".in_layers.0." in name or ".out_layers.0." in name or ".norm." in name or name.startswith("out.0."):
                t = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            elif zeroed:
                t = torch.zeros(shape)
            else:   # nn.Conv2d / nn.Linear default: U(+-1/sqrt(fan_in)) for the weight and the bias (fan_in of the weight)
                fan_in = math.prod(shapes.get(stem + ".weight", shape)[1:])
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
            sd[name] = t
        return sd

    def state_dict(self):
        if self._host is None:
            self._host = self._default_init()
        return dict(self._host)

    def load_state_dict(self, sd, strict=True):
        want = dict(self._specs)
        missing = [k for k in want if k not in sd]
        extra = [k for k in sd if k not in want]
        if strict and (missing or extra):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:4]}... unexpected {extra[:4]}...")
        host = self.state_dict()
        for k, v in sd.items():
            if k in want:
                if tuple(v.shape) != want[k]:
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {want[k]}")
                host[k] = v.detach().to("cpu", torch.float32).contiguous()
        self._host = host
        self._uploaded = False
        if self.device is not None:
            self._upload()
        return self

    def _upload(self):
        _lib.require_cuda()
        with torch.cuda.device(self.device):
            for name, t in self.state_dict().items():
                t = t.contiguous()
                _lib.check(self._L.osm_unet_load_param(self._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel(),
                                                       _lib.stream()))
            torch.cuda.synchronize()
        self._uploaded = True

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.OsmError("the native UNet runs on CUDA devices only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self._upload()
        return self

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def eval(self):
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("the native engine is inference + input-gradient only")
        return self

    def parameters(self):
        return iter(())

    def requires_grad_(self, flag=False):
        return self

    # ---- execution --------------------------------------------------------------------------------
    def _ensure_bound(self, B, H, W):
        if self.device is None or not self._uploaded:
            raise _lib.OsmError("call model.to('cuda') before running the UNet")
        if self._bound == (B, H, W):
            return
        nbytes = self._L.osm_unet_workspace_bytes(self._h, B, H, W)
        if nbytes < 0:
            _lib.check(-1)
        self._ws = None
        self._ws = torch.empty(nbytes + 512, dtype=torch.uint8, device=self.device)
        base = (self._ws.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(self.device):
            _lib.check(self._L.osm_unet_bind(self._h, B, H, W, C.c_void_p(base), nbytes))
        self._bound = (B, H, W)
        self._bind_generation += 1

    def workspace_bytes(self, B, H, W):
        """Device bytes a bind at (B, H, W) needs.  Sizing re-plans the engine (the C call drops its current binding), so the
        size of the live binding is answered from the Python side and any other query marks the model as unbound."""
        if self._bound == (B, H, W) and self._ws is not None:
            return int(self._ws.numel()) - 512
        n = int(self._L.osm_unet_workspace_bytes(self._h, B, H, W))
        self._bound = None
        return n

    def launch_counts(self):
        return self._L.osm_unet_launch_count(self._h, 0), self._L.osm_unet_launch_count(self._h, 1)

    def forward_flops(self):
        return float(self._L.osm_unet_forward_flops(self._h))

    def profile_ops(self, which):
        """Per-op device times of the forward (0) / input-VJP (1) program: list of dicts (kind, ms, flops, bytes, dims)."""
        cap = 4096
        ms, kinds = (C.c_float * cap)(), (C.c_int * cap)()
        fl, by, dims = (C.c_double * cap)(), (C.c_double * cap)(), (C.c_int * (6 * cap))()
        n = self._L.osm_unet_profile_ops(self._h, which, _lib.stream(), cap, ms, kinds, fl, by, dims)
        if n < 0:
            _lib.check(n)
        names = ["conv", "gn_stats", "gn_apply", "gn_bwd", "attn_fwd", "attn_bwd", "linear", "gn_final", "gn_coef", "gn_coef", "gn_coef", "gn_final"]
        return [dict(kind=names[kinds[i]], ms=float(ms[i]), flops=float(fl[i]), bytes=float(by[i]),
                     dims=[int(dims[6 * i + k]) for k in range(6)]) for i in range(n)]

    def _forward_raw(self, x, t, out=None):
        B, Cin, H, W = x.shape
        assert Cin == self.in_channels
        self._ensure_bound(B, H, W)
        x = x.detach().contiguous().float()
        t = t.detach().to(torch.float32).contiguous()
        if out is None:
            out = torch.empty(B, self.out_channels, H, W, dtype=torch.float32, device=x.device)
        _lib.check(self._L.osm_unet_forward(self._h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), _lib.stream()))
        self._generation += 1
        return out

    def _vjp_raw(self, grad_out, grad_x=None):
        B, H, W = self._bound
        if grad_x is None:
            grad_x = torch.empty(B, self.in_channels, H, W, dtype=torch.float32, device=grad_out.device)
        _lib.check(self._L.osm_unet_vjp_input(self._h, _lib.ptr(grad_out), _lib.ptr(grad_x), _lib.stream()))
        return grad_x

    def __call__(self, x, timesteps, y=None):
        assert y is None, "class conditioning is not supported"
        if not x.is_cuda:
            raise _lib.OsmError("the native UNet needs CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and x.requires_grad:
            return _UNetFn.apply(x, timesteps, self)
        return self._forward_raw(x, timesteps)

    forward = __call__

    def __del__(self):
        try:
            self._L.osm_unet_destroy(self._h)
        except Exception:
            pass
