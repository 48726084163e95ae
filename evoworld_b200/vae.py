"""AutoencoderKLTemporalDecoder on sm_100a (the VAE around the denoise loop).

Mirrors the diffusers 0.31 class the reference pipeline holds as `self.vae`
(evoworld/pipeline/pipeline_evoworld.py:307-328 `_encode_vae_image`, :358-385 `decode_latents`, call sites :606-623, :731):
same constructor arguments, `.config` (scaling_factor, force_upcast, ...), `from_pretrained(path, subfolder="vae")`,
diffusers state-dict key layout, `encode(x).latent_dist.mode()/.sample()`, `decode(z, num_frames=n).sample` and
`forward(...)` whose signature carries `num_frames` (decode_latents inspects it, :364-365).  The arithmetic runs in the
C-ABI library (`evw_vae_encode` / `evw_vae_decode`, csrc/vae_host.cu); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from collections import OrderedDict
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from . import _lib, ops

DEFAULT_CONFIG = dict(
    in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
    layers_per_block=2, latent_channels=4, sample_size=768, scaling_factor=0.18215, force_upcast=True,
)

MAX_FRAMES_PER_CALL = 8  # frames per evw_vae_encode / evw_vae_decode call (bounds the workspace: ~4 GB per frame at 576x1024)


class DiagonalGaussianDistribution:
    """diffusers models/autoencoders/vae.py: parameters = cat(mean, logvar) along channels, logvar clamped to [-30, 20]."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution = None


@dataclass
class DecoderOutput:
    sample: torch.Tensor = None


# ---------------------------------------------------------------------------------------------
# parameter specification: diffusers key layout
# ---------------------------------------------------------------------------------------------


def _resnet2d_spec(spec, p, cin, cout):
    spec[p + ".norm1.weight"] = (cin,); spec[p + ".norm1.bias"] = (cin,)
    spec[p + ".conv1.weight"] = (cout, cin, 3, 3); spec[p + ".conv1.bias"] = (cout,)
    spec[p + ".norm2.weight"] = (cout,); spec[p + ".norm2.bias"] = (cout,)
    spec[p + ".conv2.weight"] = (cout, cout, 3, 3); spec[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        spec[p + ".conv_shortcut.weight"] = (cout, cin, 1, 1); spec[p + ".conv_shortcut.bias"] = (cout,)


def _st_res_spec(spec, p, cin, cout):
    _resnet2d_spec(spec, p + ".spatial_res_block", cin, cout)
    t = p + ".temporal_res_block"
    spec[t + ".norm1.weight"] = (cout,); spec[t + ".norm1.bias"] = (cout,)
    spec[t + ".conv1.weight"] = (cout, cout, 3, 1, 1); spec[t + ".conv1.bias"] = (cout,)
    spec[t + ".norm2.weight"] = (cout,); spec[t + ".norm2.bias"] = (cout,)
    spec[t + ".conv2.weight"] = (cout, cout, 3, 1, 1); spec[t + ".conv2.bias"] = (cout,)
    spec[p + ".time_mixer.mix_factor"] = (1,)


def _attn_spec(spec, p, c):
    spec[p + ".group_norm.weight"] = (c,); spec[p + ".group_norm.bias"] = (c,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        spec[f"{p}.{n}.weight"] = (c, c); spec[f"{p}.{n}.bias"] = (c,)


def layout(cfg) -> dict:
    """Walk of the architecture: (prefix, cin, cout) of every res block, attention prefixes, resampling convolutions."""
    boc, lpb = tuple(cfg["block_out_channels"]), cfg["layers_per_block"]
    enc_res, dec_res, att, samplers = [], [], [], []
    prev = boc[0]
    for i, ch in enumerate(boc):
        for j in range(lpb):
            enc_res.append((f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else ch, ch))
        if i < len(boc) - 1:
            samplers.append((f"encoder.down_blocks.{i}.downsamplers.0.conv", ch))
        prev = ch
    enc_res += [("encoder.mid_block.resnets.0", boc[-1], boc[-1]), ("encoder.mid_block.resnets.1", boc[-1], boc[-1])]
    att.append(("encoder.mid_block.attentions.0", boc[-1]))
    for i in range(lpb):
        dec_res.append((f"decoder.mid_block.resnets.{i}", boc[-1], boc[-1]))
    for i in range(lpb - 1):
        att.append((f"decoder.mid_block.attentions.{i}", boc[-1]))
    rev, ch = boc[::-1], boc[-1]
    for i, out_ch in enumerate(rev):
        for j in range(lpb + 1):
            dec_res.append((f"decoder.up_blocks.{i}.resnets.{j}", ch if j == 0 else out_ch, out_ch))
        if i < len(rev) - 1:
            samplers.append((f"decoder.up_blocks.{i}.upsamplers.0.conv", out_ch))
        ch = out_ch
    return dict(enc_res=enc_res, dec_res=dec_res, att=att, samplers=samplers, boc=boc)


def param_spec(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    lay, boc, lat = layout(cfg), tuple(cfg["block_out_channels"]), cfg["latent_channels"]
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    spec["encoder.conv_in.weight"] = (boc[0], cfg["in_channels"], 3, 3); spec["encoder.conv_in.bias"] = (boc[0],)
    for p, cin, cout in lay["enc_res"]:
        _resnet2d_spec(spec, p, cin, cout)
    spec["encoder.conv_norm_out.weight"] = (boc[-1],); spec["encoder.conv_norm_out.bias"] = (boc[-1],)
    spec["encoder.conv_out.weight"] = (2 * lat, boc[-1], 3, 3); spec["encoder.conv_out.bias"] = (2 * lat,)
    spec["quant_conv.weight"] = (2 * lat, 2 * lat, 1, 1); spec["quant_conv.bias"] = (2 * lat,)
    spec["decoder.conv_in.weight"] = (boc[-1], lat, 3, 3); spec["decoder.conv_in.bias"] = (boc[-1],)
    for p, cin, cout in lay["dec_res"]:
        _st_res_spec(spec, p, cin, cout)
    for p, c in lay["att"]:
        _attn_spec(spec, p, c)
    for p, c in lay["samplers"]:
        spec[p + ".weight"] = (c, c, 3, 3); spec[p + ".bias"] = (c,)
    spec["decoder.conv_norm_out.weight"] = (boc[0],); spec["decoder.conv_norm_out.bias"] = (boc[0],)
    spec["decoder.conv_out.weight"] = (cfg["out_channels"], boc[0], 3, 3); spec["decoder.conv_out.bias"] = (cfg["out_channels"],)
    spec["decoder.time_conv_out.weight"] = (cfg["out_channels"], cfg["out_channels"], 3, 1, 1)
    spec["decoder.time_conv_out.bias"] = (cfg["out_channels"],)
    return spec


class AutoencoderKLTemporalDecoder:
    """KL VAE with the temporal decoder of Stable Video Diffusion: images [N,3,H,W] in [-1,1] <-> latents [N,4,H/8,W/8]."""

    config_name = "config.json"

    def __init__(self, **kwargs):
        cfg = dict(DEFAULT_CONFIG)
        unknown = set(kwargs) - set(cfg) - {"_class_name", "_diffusers_version", "_name_or_path"}
        if unknown:
            raise TypeError(f"unexpected config keys: {sorted(unknown)}")
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        if len(cfg["block_out_channels"]) != 4 or len(cfg["down_block_types"]) != 4:
            raise NotImplementedError("evoworld_b200 builds the 4-level VAE Stable Video Diffusion uses")
        if cfg["in_channels"] != 3 or cfg["out_channels"] != 3:
            raise NotImplementedError("RGB images only")
        self._cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self._spec = param_spec(cfg)
        self._params: Dict[str, torch.Tensor] = {}
        self._device = torch.device("cpu")
        self._handle = None
        self._packed = None
        self._ws: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ parameters
    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return torch.float32

    def num_parameters(self) -> int:
        return sum(math.prod(s) for s in self._spec.values())

    def init_random(self, seed: int = 0, device=None):
        """PyTorch-default initialisation drawn directly on `device` (norms 1/0, mix_factor 0 as diffusers' merge_factor)."""
        dev = torch.device(device) if device is not None else self._device
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        self._params = {}
        for name, shape in self._spec.items():
            if name.endswith("mix_factor"):
                t = torch.zeros(shape, device=dev)
            elif "norm" in name:
                t = torch.ones(shape, device=dev) if name.endswith("weight") else torch.zeros(shape, device=dev)
            else:
                wshape = self._spec[name.rsplit(".", 1)[0] + ".weight"]
                bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
                t = (torch.rand(shape, device=dev, generator=g) * 2 - 1) * bound
            self._params[name] = t
        self._device = dev
        self._invalidate()
        return self

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, self._params[k]) for k in self._spec)

    def load_state_dict(self, sd, strict: bool = True):
        missing = [k for k in self._spec if k not in sd]
        unexpected = [k for k in sd if k not in self._spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shape in self._spec.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {shape}")
                self._params[k] = sd[k].detach().to(self._device, torch.float32, copy=True).contiguous()
        self._invalidate()
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def parameters(self):
        return iter(self._params.values())

    def named_parameters(self):
        return iter(self._params.items())

    def requires_grad_(self, flag: bool = False):
        return self

    def eval(self):
        return self

    def to(self, device=None, dtype=None, **_):
        """Moves the parameters; a dtype is accepted and ignored (the pipeline toggles the VAE between fp16 and fp32 around
        `force_upcast`, pipeline_evoworld.py:606-623 — the kernels always take fp16 operands with fp32 accumulation)."""
        if isinstance(device, torch.dtype):
            device, dtype = None, device
        if device is not None and torch.device(device) != self._device:
            self._device = torch.device(device)
            self._params = {k: v.to(self._device) for k, v in self._params.items()}
            self._invalidate()
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    @classmethod
    def from_config(cls, config: dict):
        return cls(**{k: v for k, v in config.items() if k in DEFAULT_CONFIG})

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, torch_dtype=None,
                        variant: Optional[str] = None, **kwargs):
        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        cfg_path = os.path.join(root, cls.config_name)
        if not os.path.isfile(cfg_path):
            raise FileNotFoundError(f"{cfg_path} not found")
        with open(cfg_path) as f:
            model = cls.from_config(json.load(f))
        stem = "diffusion_pytorch_model" + (f".{variant}" if variant else "")
        st_path, bin_path = os.path.join(root, stem + ".safetensors"), os.path.join(root, stem + ".bin")
        if os.path.isfile(st_path):
            from safetensors.torch import load_file

            sd = load_file(st_path)
        elif os.path.isfile(bin_path):
            sd = torch.load(bin_path, map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights found under {root}")
        model.load_state_dict(sd)
        return model

    def save_pretrained(self, save_directory: str):
        from safetensors.torch import save_file

        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump({**self._cfg, "_class_name": "AutoencoderKLTemporalDecoder"}, f, indent=2, default=list)
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()},
                  os.path.join(save_directory, "diffusion_pytorch_model.safetensors"))

    # ------------------------------------------------------------------ packing + handle
    def _invalidate(self):
        if self._handle is not None:
            _lib.lib().evw_vae_destroy(self._handle)
        self._handle, self._packed, self._ws = None, None, None

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().evw_vae_destroy(self._handle)
        except Exception:
            pass

    @torch.no_grad()
    def pack_parameters(self):
        """fp32 diffusers-layout parameters -> kernel layouts.  Returns (tensors: name -> CUDA tensor, scalars: name -> float)."""
        P, cfg = self._params, self._cfg
        if self._device.type != "cuda":
            raise RuntimeError("evoworld_b200 VAE: parameters must be on a CUDA device (no CPU fallback)")
        if len(P) != len(self._spec):
            raise RuntimeError("VAE parameters are not initialised (load_state_dict / init_random first)")
        lay = layout(cfg)
        T: Dict[str, torch.Tensor] = {}
        S: Dict[str, float] = {}
        h = lambda t: t.to(torch.float16).contiguous()
        f = lambda t: t.to(torch.float32).contiguous()

        def conv2d_w(w, pad_out=0):  # [O,I,3,3] -> [O, ky, kx, I] -> [O, 9 I]
            if pad_out:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, pad_out))
            return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)

        def conv_in_split(name, cin):  # per tap [W_hi | W_hi | W_lo | 0] against the operand channels [head | tail | head | 0]
            w = P[name + ".weight"].permute(0, 2, 3, 1).to(torch.float32)  # [O, ky, kx, I]
            hi = w.to(torch.float16)
            lo = (w - hi.to(torch.float32)).to(torch.float16)
            wk = torch.cat([hi, hi, lo, hi.new_zeros(*hi.shape[:-1], 64 - 3 * cin)], dim=-1)
            T[name + ".weight"] = wk.reshape(wk.shape[0], -1).contiguous(); T[name + ".bias"] = f(P[name + ".bias"])

        def norm(name):
            T[name + ".weight"] = f(P[name + ".weight"]); T[name + ".bias"] = f(P[name + ".bias"])

        def resnet2d(p, cin, cout):
            norm(p + ".norm1"); norm(p + ".norm2")
            T[p + ".conv1.weight"] = h(conv2d_w(P[p + ".conv1.weight"])); T[p + ".conv1.bias"] = f(P[p + ".conv1.bias"])
            w2, b2 = conv2d_w(P[p + ".conv2.weight"]), P[p + ".conv2.bias"]
            if cin != cout:  # the 1x1 shortcut of the raw input rides along as a tenth tap
                w2 = torch.cat([w2, P[p + ".conv_shortcut.weight"].reshape(cout, cin)], dim=1)
                b2 = b2 + P[p + ".conv_shortcut.bias"]
            T[p + ".conv2.weight"] = h(w2); T[p + ".conv2.bias"] = f(b2)

        conv_in_split("encoder.conv_in", cfg["in_channels"])
        conv_in_split("decoder.conv_in", cfg["latent_channels"])
        for p, cin, cout in lay["enc_res"]:
            resnet2d(p, cin, cout)
        for p, cin, cout in lay["dec_res"]:
            resnet2d(p + ".spatial_res_block", cin, cout)
            t = p + ".temporal_res_block"
            norm(t + ".norm1"); norm(t + ".norm2")
            for cname in (".conv1", ".conv2"):  # [O,I,3,1,1] -> [O, kt, I]
                T[t + cname + ".weight"] = h(P[t + cname + ".weight"][..., 0, 0].permute(0, 2, 1).reshape(cout, -1))
                T[t + cname + ".bias"] = f(P[t + cname + ".bias"])
            # AlphaBlender(merge_strategy="learned", switch_spatial_to_temporal_mix=True): weight of the spatial branch
            S[p + ".time_mixer.alpha"] = 1.0 - float(torch.sigmoid(P[p + ".time_mixer.mix_factor"]).item())
        for p, c in lay["att"]:
            norm(p + ".group_norm")
            for n in ("to_q", "to_k"):
                T[f"{p}.{n}.weight"] = h(P[f"{p}.{n}.weight"]); T[f"{p}.{n}.bias"] = f(P[f"{p}.{n}.bias"])
            T[p + ".to_v.weight"] = h(P[p + ".to_v.weight"])
            # softmax rows sum to one: P (V + 1 b_v^T) = P V + b_v^T, so to_v's bias moves into to_out's
            wo = P[p + ".to_out.0.weight"]
            T[p + ".to_out.0.weight"] = h(wo)
            T[p + ".to_out.0.bias"] = f(P[p + ".to_out.0.bias"].double() + wo.double() @ P[p + ".to_v.bias"].double())
        for p, c in lay["samplers"]:
            T[p + ".weight"] = h(conv2d_w(P[p + ".weight"])); T[p + ".bias"] = f(P[p + ".bias"])
            if ".upsamplers." in p:  # nearest x2 + 3x3 conv fused into four 2x2 phase convolutions (ops.upconv_weights)
                T[p + ".weight4"] = ops.upconv_weights(P[p + ".weight"])
        norm("encoder.conv_norm_out"); norm("decoder.conv_norm_out")
        # quant_conv (1x1) composed with encoder.conv_out: W' = Wq Wc, b' = Wq bc + bq
        lat2 = 2 * cfg["latent_channels"]
        wq = P["quant_conv.weight"].reshape(lat2, lat2).double()
        wc = conv2d_w(P["encoder.conv_out.weight"]).double()
        T["encoder.conv_out.weight"] = h(torch.nn.functional.pad((wq @ wc).float(), (0, 0, 0, 16 - lat2)))
        T["encoder.conv_out.bias"] = f(torch.nn.functional.pad((wq @ P["encoder.conv_out.bias"].double() + P["quant_conv.bias"].double()).float(),
                                                              (0, 16 - lat2)))
        co = cfg["out_channels"]
        T["decoder.conv_out.weight"] = h(conv2d_w(P["decoder.conv_out.weight"], pad_out=16 - co))
        T["decoder.conv_out.bias"] = f(torch.nn.functional.pad(P["decoder.conv_out.bias"], (0, 16 - co)))
        T["decoder.time_conv_out.weight"] = f(P["decoder.time_conv_out.weight"][..., 0, 0])  # [o, i, kt]
        T["decoder.time_conv_out.bias"] = f(P["decoder.time_conv_out.bias"])
        return T, S

    def _ensure_handle(self):
        if self._handle is not None:
            return
        L = _lib.lib()
        tensors, scalars = self.pack_parameters()
        cfg = self._cfg
        ints = [cfg["in_channels"], cfg["out_channels"], cfg["latent_channels"], *cfg["block_out_channels"], cfg["layers_per_block"]]
        names = list(tensors)
        tn = (C.c_char_p * len(names))(*[n.encode() for n in names])
        tp = (C.c_void_p * len(names))(*[tensors[n].data_ptr() for n in names])
        sn_list = list(scalars)
        sn = (C.c_char_p * len(sn_list))(*[n.encode() for n in sn_list])
        sv = (C.c_double * len(sn_list))(*[scalars[n] for n in sn_list])
        handle = C.c_void_p()
        _lib.check(L.evw_vae_create(C.byref(handle), (C.c_int * len(ints))(*ints), len(ints), tn, tp, len(names), sn, sv,
                                    len(sn_list)), "evw_vae_create")
        self._handle, self._packed = handle, tensors

    def free_master_parameters(self):
        self._ensure_handle()
        self._params = {}

    def _workspace(self, mode: int, n: int, num_frames: int, H: int, W: int) -> torch.Tensor:
        need = _lib.lib().evw_vae_workspace_bytes(self._handle, mode, n, num_frames, H, W)
        if need < 0:
            _lib.check(-1, "evw_vae_workspace_bytes")
        if self._ws is None or self._ws.numel() < need + 1024:
            self._ws = None  # release before growing
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self._device)
        off = (-self._ws.data_ptr()) % 1024
        return self._ws[off:off + need]

    def plan_info(self, mode: int):
        """(kernel launches, algorithmic FLOPs, GroupNorms fed by GEMM epilogues) of the last encode (0) / decode (1) plan."""
        a, b, c = C.c_int64(), C.c_double(), C.c_int64()
        _lib.check(_lib.lib().evw_vae_plan_info(self._handle, mode, C.byref(a), C.byref(b), C.byref(c)), "evw_vae_plan_info")
        return a.value, b.value, c.value

    # ------------------------------------------------------------------ compute
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [N,3,H,W] in [-1,1] -> AutoencoderKLOutput(latent_dist) (pipeline_evoworld.py:315 takes `.latent_dist.mode()`)."""
        _lib.require_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != self._cfg["in_channels"]:
            raise ValueError(f"encode expects [N,{self._cfg['in_channels']},H,W], got {tuple(x.shape)}")
        self._ensure_handle()
        N, _, H, W = x.shape
        xin = x.to(torch.float32).contiguous()
        lat2 = 2 * self._cfg["latent_channels"]
        moments = torch.empty((N, lat2, H // 8, W // 8), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            for i in range(0, N, MAX_FRAMES_PER_CALL):
                n = min(MAX_FRAMES_PER_CALL, N - i)
                ws = self._workspace(0, n, 1, H, W)
                _lib.check(_lib.lib().evw_vae_encode(self._handle, _lib.ptr(xin[i:i + n]), _lib.ptr(moments[i:i + n]), n, H, W,
                                                     ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device)), "evw_vae_encode")
        dist = DiagonalGaussianDistribution(moments.to(x.dtype))
        return AutoencoderKLOutput(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, num_frames: int = 1, return_dict: bool = True):
        """z [N,4,h,w] (N = videos x num_frames; already divided by scaling_factor, pipeline_evoworld.py:362) ->
        DecoderOutput(sample [N,3,8h,8w])."""
        _lib.require_cuda(z, "z")
        if z.dim() != 4 or z.shape[1] != self._cfg["latent_channels"]:
            raise ValueError(f"decode expects [N,{self._cfg['latent_channels']},h,w], got {tuple(z.shape)}")
        N, _, h, w = z.shape
        if num_frames < 1 or N % num_frames:
            raise ValueError(f"{N} latents are not a multiple of num_frames={num_frames}")
        self._ensure_handle()
        zin = z.to(torch.float32).contiguous()
        out = torch.empty((N, self._cfg["out_channels"], 8 * h, 8 * w), dtype=torch.float32, device=z.device)
        # whole videos per call (the temporal blocks and time_conv_out mix the frames of a video)
        per_call = max(1, MAX_FRAMES_PER_CALL // num_frames) * num_frames
        with torch.cuda.device(z.device):
            for i in range(0, N, per_call):
                n = min(per_call, N - i)
                ws = self._workspace(1, n, num_frames, h, w)
                _lib.check(_lib.lib().evw_vae_decode(self._handle, _lib.ptr(zin[i:i + n]), _lib.ptr(out[i:i + n]), n, num_frames, h, w,
                                                     ws.data_ptr(), ws.numel(), _lib.stream_ptr(z.device)), "evw_vae_decode")
        out = out.to(z.dtype)
        return DecoderOutput(sample=out) if return_dict else (out,)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None, num_frames: int = 1):
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z, num_frames=num_frames).sample
        return DecoderOutput(sample=dec) if return_dict else (dec,)

    __call__ = forward
