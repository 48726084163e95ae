"""UNetSpatioTemporalConditionModel on sm_100a (operator boundary 1).

Mirrors evoworld/trainer/unet_plucker.py:30-488 of the reference: same constructor arguments,
`.config`, `from_pretrained(path, subfolder="unet")`, diffusers state-dict key layout,
and `forward(sample, timestep, encoder_hidden_states, added_time_ids, return_dict)`.  The arithmetic
runs in the C-ABI library: parameters are packed once into the kernel layouts (fp16 tap-major
convolution / linear weights, fused qkv, interleaved GEGLU, folded single-key cross-attention) and
`evw_unet_forward` / `evw_denoise_step` enqueue the whole network.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from collections import OrderedDict
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional, Tuple, Union

import torch

from . import _lib, ops

MAX_FRAMES = 32


@dataclass
class UNetSpatioTemporalConditionOutput:
    sample: torch.Tensor = None


DEFAULT_CONFIG = dict(
    sample_size=None, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                      "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                    "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256, projection_class_embeddings_input_dim=768,
    layers_per_block=2, cross_attention_dim=1024, transformer_layers_per_block=1, num_attention_heads=(5, 10, 20, 20),
    num_frames=25,
)


# ---------------------------------------------------------------------------------------------
# parameter specification: diffusers key layout
# ---------------------------------------------------------------------------------------------


def _resblock_spec(spec, p, cin, cout, temb):
    s, t = p + ".spatial_res_block", p + ".temporal_res_block"
    spec[s + ".norm1.weight"] = (cin,); spec[s + ".norm1.bias"] = (cin,)
    spec[s + ".conv1.weight"] = (cout, cin, 3, 3); spec[s + ".conv1.bias"] = (cout,)
    spec[s + ".time_emb_proj.weight"] = (cout, temb); spec[s + ".time_emb_proj.bias"] = (cout,)
    spec[s + ".norm2.weight"] = (cout,); spec[s + ".norm2.bias"] = (cout,)
    spec[s + ".conv2.weight"] = (cout, cout, 3, 3); spec[s + ".conv2.bias"] = (cout,)
    if cin != cout:
        spec[s + ".conv_shortcut.weight"] = (cout, cin, 1, 1); spec[s + ".conv_shortcut.bias"] = (cout,)
    spec[t + ".norm1.weight"] = (cout,); spec[t + ".norm1.bias"] = (cout,)
    spec[t + ".conv1.weight"] = (cout, cout, 3, 1, 1); spec[t + ".conv1.bias"] = (cout,)
    spec[t + ".time_emb_proj.weight"] = (cout, temb); spec[t + ".time_emb_proj.bias"] = (cout,)
    spec[t + ".norm2.weight"] = (cout,); spec[t + ".norm2.bias"] = (cout,)
    spec[t + ".conv2.weight"] = (cout, cout, 3, 1, 1); spec[t + ".conv2.bias"] = (cout,)
    spec[p + ".time_mixer.mix_factor"] = (1,)


def _attn_spec(spec, p, dim, kv_dim):
    spec[p + ".to_q.weight"] = (dim, dim)
    spec[p + ".to_k.weight"] = (dim, kv_dim)
    spec[p + ".to_v.weight"] = (dim, kv_dim)
    spec[p + ".to_out.0.weight"] = (dim, dim); spec[p + ".to_out.0.bias"] = (dim,)


def _ff_spec(spec, p, dim):
    spec[p + ".net.0.proj.weight"] = (8 * dim, dim); spec[p + ".net.0.proj.bias"] = (8 * dim,)
    spec[p + ".net.2.weight"] = (dim, 4 * dim); spec[p + ".net.2.bias"] = (dim,)


def _transformer_spec(spec, p, c, cross):
    spec[p + ".norm.weight"] = (c,); spec[p + ".norm.bias"] = (c,)
    spec[p + ".proj_in.weight"] = (c, c); spec[p + ".proj_in.bias"] = (c,)
    b = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        spec[f"{b}.{n}.weight"] = (c,); spec[f"{b}.{n}.bias"] = (c,)
    _attn_spec(spec, b + ".attn1", c, c)
    _attn_spec(spec, b + ".attn2", c, cross)
    _ff_spec(spec, b + ".ff", c)
    t = p + ".temporal_transformer_blocks.0"
    for n in ("norm_in", "norm1", "norm2", "norm3"):
        spec[f"{t}.{n}.weight"] = (c,); spec[f"{t}.{n}.bias"] = (c,)
    _ff_spec(spec, t + ".ff_in", c)
    _attn_spec(spec, t + ".attn1", c, c)
    _attn_spec(spec, t + ".attn2", c, cross)
    _ff_spec(spec, t + ".ff", c)
    spec[p + ".time_pos_embed.linear_1.weight"] = (4 * c, c); spec[p + ".time_pos_embed.linear_1.bias"] = (4 * c,)
    spec[p + ".time_pos_embed.linear_2.weight"] = (c, 4 * c); spec[p + ".time_pos_embed.linear_2.bias"] = (c,)
    spec[p + ".time_mixer.mix_factor"] = (1,)
    spec[p + ".proj_out.weight"] = (c, c); spec[p + ".proj_out.bias"] = (c,)


def block_layout(cfg) -> dict:
    """Walk of the architecture (unet_plucker.py:163-233): res blocks and transformers with channels."""
    boc = tuple(cfg["block_out_channels"])
    heads = cfg["num_attention_heads"]
    heads = (heads,) * len(boc) if isinstance(heads, int) else tuple(heads)
    lpb = cfg["layers_per_block"]
    down_attn = tuple(t.startswith("CrossAttn") for t in cfg["down_block_types"])
    res, att, samplers = [], [], []
    out_ch = boc[0]
    for i in range(len(boc)):
        in_ch, out_ch = out_ch, boc[i]
        for j in range(lpb):
            res.append((f"down_blocks.{i}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch))
            if down_attn[i]:
                att.append((f"down_blocks.{i}.attentions.{j}", out_ch, heads[i]))
        if i < len(boc) - 1:
            samplers.append((f"down_blocks.{i}.downsamplers.0.conv", out_ch))
    res.append(("mid_block.resnets.0", boc[-1], boc[-1]))
    att.append(("mid_block.attentions.0", boc[-1], heads[-1]))
    res.append(("mid_block.resnets.1", boc[-1], boc[-1]))
    rev, rev_heads, rev_attn = boc[::-1], heads[::-1], down_attn[::-1]
    out_ch = rev[0]
    for i in range(len(boc)):
        prev, out_ch = out_ch, rev[i]
        in_ch = rev[min(i + 1, len(boc) - 1)]
        for j in range(lpb + 1):
            skip = in_ch if j == lpb else out_ch
            rin = prev if j == 0 else out_ch
            res.append((f"up_blocks.{i}.resnets.{j}", rin + skip, out_ch))
            if rev_attn[i]:
                att.append((f"up_blocks.{i}.attentions.{j}", out_ch, rev_heads[i]))
        if i < len(boc) - 1:
            samplers.append((f"up_blocks.{i}.upsamplers.0.conv", out_ch))
    return dict(res=res, att=att, samplers=samplers, boc=boc, heads=heads, down_attn=down_attn)


def param_spec(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    lay = block_layout(cfg)
    boc = lay["boc"]
    temb = boc[0] * 4
    cross = cfg["cross_attention_dim"]
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    spec["conv_in.weight"] = (boc[0], cfg["in_channels"], 3, 3); spec["conv_in.bias"] = (boc[0],)
    for name, din in (("time_embedding", boc[0]), ("add_embedding", cfg["projection_class_embeddings_input_dim"])):
        spec[f"{name}.linear_1.weight"] = (temb, din); spec[f"{name}.linear_1.bias"] = (temb,)
        spec[f"{name}.linear_2.weight"] = (temb, temb); spec[f"{name}.linear_2.bias"] = (temb,)
    for p, cin, cout in lay["res"]:
        _resblock_spec(spec, p, cin, cout, temb)
    for p, c, _ in lay["att"]:
        _transformer_spec(spec, p, c, cross)
    for p, c in lay["samplers"]:
        spec[p + ".weight"] = (c, c, 3, 3); spec[p + ".bias"] = (c,)
    spec["conv_norm_out.weight"] = (boc[0],); spec["conv_norm_out.bias"] = (boc[0],)
    spec["conv_out.weight"] = (cfg["out_channels"], boc[0], 3, 3); spec["conv_out.bias"] = (cfg["out_channels"],)
    return spec


def _sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t[:, None].float() * freq[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


def _geglu_interleave(w, b):
    F2, K = w.shape
    F = F2 // 2
    wi = torch.stack([w[:F].reshape(F // 16, 16, K), w[F:].reshape(F // 16, 16, K)], dim=1).reshape(F2, K)
    bi = torch.stack([b[:F].reshape(F // 16, 16), b[F:].reshape(F // 16, 16)], dim=1).reshape(F2)
    return wi, bi


class _LinearInfo:
    def __init__(self, in_features, out_features):
        self.in_features, self.out_features = in_features, out_features


class UNetSpatioTemporalConditionModel:
    """Conditional spatio-temporal UNet (sample [B,T,C,h,w] -> [B,T,4,h,w]); EvoWorld widens conv_in to
    18 = 4 (noisy) + 4 (first frame) + 4 (memory) + 6 (Plücker) channels (trainer_utils.py:19)."""

    config_name = "config.json"

    def __init__(self, **kwargs):
        cfg = dict(DEFAULT_CONFIG)
        unknown = set(kwargs) - set(cfg) - {"_class_name", "_diffusers_version", "_name_or_path"}
        if unknown:
            raise TypeError(f"unexpected config keys: {sorted(unknown)}")
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        if len(cfg["down_block_types"]) != len(cfg["up_block_types"]):
            raise ValueError("Must provide the same number of `down_block_types` as `up_block_types`.")
        if len(cfg["block_out_channels"]) != len(cfg["down_block_types"]):
            raise ValueError("Must provide the same number of `block_out_channels` as `down_block_types`.")
        if len(cfg["block_out_channels"]) != 4:
            raise NotImplementedError("evoworld_b200 builds the 4-level UNet EvoWorld uses")
        heads = cfg["num_attention_heads"]
        if not isinstance(heads, int) and len(heads) != len(cfg["down_block_types"]):
            raise ValueError("Must provide the same number of `num_attention_heads` as `down_block_types`.")
        if cfg["transformer_layers_per_block"] not in (1, [1, 1, 1, 1], (1, 1, 1, 1)):
            raise NotImplementedError("transformer_layers_per_block != 1")
        self._cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self._spec = param_spec(cfg)
        self._params: Dict[str, torch.Tensor] = {}
        self._device = torch.device("cpu")
        self._handle = None
        self._packed = None
        self._ws: Dict[Tuple[int, int, int, int], torch.Tensor] = {}
        self.add_embedding = SimpleNamespace(
            linear_1=_LinearInfo(cfg["projection_class_embeddings_input_dim"], cfg["block_out_channels"][0] * 4))
        self.eps = dict(cross=1e-6, plain=1e-5, mid=1e-5, up=1e-6)  # diffusers block defaults
        # Split precision (default on; EVW_UNET_SPLIT=0 for A/B runs): conv_in, conv_out, proj_in, proj_out take fp16
        # head + tail operands and the spatial conv1 output stays fp32 — the layers tools/precision_sim.py ranks as
        # 3/4 of the fp16 rounding-error variance of one forward for < 4 % of its FLOPs (1.1e-3 -> 8e-4 rel. L2).
        self.precision_split = os.environ.get("EVW_UNET_SPLIT", "1") != "0" and 3 * cfg["in_channels"] <= 64 \
            and 2 * cfg["out_channels"] <= 16

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
        """PyTorch-default initialisation (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases,
        norms 1/0, mix_factor 0.5) drawn directly on `device`."""
        dev = torch.device(device) if device is not None else self._device
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        self._params = {}
        for name, shape in self._spec.items():
            if name.endswith("mix_factor"):
                t = torch.full(shape, 0.5, device=dev)
            elif ".norm" in name or name.startswith("conv_norm_out"):
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
                # always a private, freshly allocated (hence 512-byte aligned) copy: the kernels read parameters with
                # 16-byte vector loads, and a caller's tensors may be views at arbitrary offsets of one flat buffer
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
        """Load a diffusers-format checkpoint directory: <path>/<subfolder>/config.json +
        diffusion_pytorch_model[.variant].safetensors (or .bin)."""
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
            json.dump({**self._cfg, "_class_name": "UNetSpatioTemporalConditionModel"}, f, indent=2, default=list)
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()},
                  os.path.join(save_directory, "diffusion_pytorch_model.safetensors"))

    # ------------------------------------------------------------------ packing + handle
    def _invalidate(self):
        if self._handle is not None:
            _lib.lib().evw_unet_destroy(self._handle)
        self._handle, self._packed, self._ws = None, None, {}

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().evw_unet_destroy(self._handle)
        except Exception:
            pass

    @torch.no_grad()
    def pack_parameters(self):
        """fp32 diffusers-layout parameters -> kernel layouts.  Returns (tensors: name -> CUDA tensor,
        scalars: name -> float, temb_total, xattn_total)."""
        P, cfg = self._params, self._cfg
        if self._device.type != "cuda":
            raise RuntimeError("evoworld_b200 UNet: parameters must be on a CUDA device (no CPU fallback)")
        if len(P) != len(self._spec):
            raise RuntimeError("UNet parameters are not initialised (load_state_dict / init_random first)")
        lay = block_layout(cfg)
        T: Dict[str, torch.Tensor] = {}
        S: Dict[str, float] = {}
        h = lambda t: t.to(torch.float16).contiguous()
        f = lambda t: t.to(torch.float32).contiguous()
        split = bool(self.precision_split)
        S["precision.split"] = 1.0 if split else 0.0

        def hi_lo(w):  # fp16 head and fp16 tail (the head's rounding error) of an fp32 weight
            w = w.to(torch.float32)
            hi = w.to(torch.float16)
            return hi, (w - hi.to(torch.float32)).to(torch.float16)

        def conv2d_w(w, pad_in=0, pad_out=0):  # [O,I,3,3] -> [O, ky, kx, I] -> [O, 9 I]
            if pad_in:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, pad_in))
            if pad_out:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, pad_out))
            return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)

        def lin(name):
            T[name + ".weight"] = h(P[name + ".weight"]); T[name + ".bias"] = f(P[name + ".bias"])

        def lin_split(name):  # [N, 3K] = [W_hi | W_hi | W_lo] against the taps [a_hi | a_lo | a_hi]
            hi, lo = hi_lo(P[name + ".weight"])
            T[name + ".weight"] = torch.cat([hi, hi, lo], dim=1).contiguous(); T[name + ".bias"] = f(P[name + ".bias"])

        def norm(name):
            T[name + ".weight"] = f(P[name + ".weight"]); T[name + ".bias"] = f(P[name + ".bias"])

        cin = cfg["in_channels"]
        if split:  # per tap: [W_hi | W_hi | W_lo | 0] against the operand channels [head | tail | head | 0]
            hi, lo = hi_lo(P["conv_in.weight"].permute(0, 2, 3, 1))  # [O, ky, kx, I]
            w = torch.cat([hi, hi, lo, torch.zeros_like(hi[..., : 64 - 3 * cin])], dim=-1)
            T["conv_in.weight"] = w.reshape(w.shape[0], -1).contiguous()
        else:
            T["conv_in.weight"] = h(conv2d_w(P["conv_in.weight"], pad_in=64 - cin))
        T["conv_in.bias"] = f(P["conv_in.bias"])
        for n in ("time_embedding.linear_1", "time_embedding.linear_2", "add_embedding.linear_1", "add_embedding.linear_2"):
            lin(n)
        temb_w, temb_b, off = [], [], 0
        for p, rin, rout in lay["res"]:
            s, t = p + ".spatial_res_block", p + ".temporal_res_block"
            norm(s + ".norm1"); norm(s + ".norm2"); norm(t + ".norm1"); norm(t + ".norm2")
            T[s + ".conv1.weight"] = h(conv2d_w(P[s + ".conv1.weight"])); T[s + ".conv1.bias"] = f(P[s + ".conv1.bias"])
            w2, b2 = conv2d_w(P[s + ".conv2.weight"]), P[s + ".conv2.bias"]
            if rin != rout:
                w2 = torch.cat([w2, P[s + ".conv_shortcut.weight"].reshape(rout, rin)], dim=1)
                b2 = b2 + P[s + ".conv_shortcut.bias"]
            T[s + ".conv2.weight"] = h(w2); T[s + ".conv2.bias"] = f(b2)
            for cname in (".conv1", ".conv2"):  # [O,I,3,1,1] -> [O, kt, I]
                T[t + cname + ".weight"] = h(P[t + cname + ".weight"][..., 0, 0].permute(0, 2, 1).reshape(rout, -1))
                T[t + cname + ".bias"] = f(P[t + cname + ".bias"])
            for blk in (s, t):
                temb_w.append(P[blk + ".time_emb_proj.weight"]); temb_b.append(P[blk + ".time_emb_proj.bias"])
                S[blk + ".time_emb_proj.offset"] = float(off)
                off += rout
            S[p + ".time_mixer.alpha"] = float(torch.sigmoid(P[p + ".time_mixer.mix_factor"]).item())
        T["temb_proj_all.weight"] = h(torch.cat(temb_w, 0)); T["temb_proj_all.bias"] = f(torch.cat(temb_b, 0))
        temb_total = off
        xw, xb, xoff = [], [], 0
        frames = torch.arange(MAX_FRAMES, device=self._device)
        for p, c, _heads in lay["att"]:
            norm(p + ".norm")
            (lin_split if split else lin)(p + ".proj_in"); (lin_split if split else lin)(p + ".proj_out")
            for b in (p + ".transformer_blocks.0", p + ".temporal_transformer_blocks.0"):
                for n in ("norm1", "norm3") + (("norm_in",) if "temporal" in b else ()):
                    norm(f"{b}.{n}")
                a1 = b + ".attn1"
                T[a1 + ".qkv.weight"] = h(torch.cat([P[a1 + ".to_q.weight"], P[a1 + ".to_k.weight"], P[a1 + ".to_v.weight"]], 0))
                lin(a1 + ".to_out.0")
                # attn2 attends to ONE key: softmax == 1, output = to_out(to_v(ehs)); fold the two linears
                a2 = b + ".attn2"
                xw.append(P[a2 + ".to_out.0.weight"].double() @ P[a2 + ".to_v.weight"].double())
                xb.append(P[a2 + ".to_out.0.bias"])
                S[a2 + ".offset"] = float(xoff)
                xoff += c
                for ffn in ((".ff",) if "temporal" not in b else (".ff_in", ".ff")):
                    wi, bi = _geglu_interleave(P[b + ffn + ".net.0.proj.weight"], P[b + ffn + ".net.0.proj.bias"])
                    T[b + ffn + ".net.0.proj.weight"] = h(wi); T[b + ffn + ".net.0.proj.bias"] = f(bi)
                    lin(b + ffn + ".net.2")
            # time_pos_embed(Timesteps(c)(arange(T))) depends on the weights only: tabulate MAX_FRAMES rows
            e = _sinusoid(frames, c)
            e = torch.nn.functional.silu(e @ P[p + ".time_pos_embed.linear_1.weight"].T + P[p + ".time_pos_embed.linear_1.bias"])
            T[p + ".time_pos_embed.table"] = f(e @ P[p + ".time_pos_embed.linear_2.weight"].T + P[p + ".time_pos_embed.linear_2.bias"])
            S[p + ".time_mixer.alpha"] = float(torch.sigmoid(P[p + ".time_mixer.mix_factor"]).item())
        T["xattn_all.weight"] = h(torch.cat(xw, 0).float()); T["xattn_all.bias"] = f(torch.cat(xb, 0))
        for p, c in lay["samplers"]:
            T[p + ".weight"] = h(conv2d_w(P[p + ".weight"])); T[p + ".bias"] = f(P[p + ".bias"])
            if ".upsamplers." in p:  # nearest x2 + 3x3 conv fused into four 2x2 phase convolutions (ops.upconv_weights)
                T[p + ".weight4"] = ops.upconv_weights(P[p + ".weight"])
        norm("conv_norm_out")
        co = cfg["out_channels"]
        if split:  # rows [0,co): [9 taps W_hi (head operand) | 9 taps W_hi (tail operand)]; rows [co,2co): [W_lo | 0]
            hi, lo = hi_lo(conv2d_w(P["conv_out.weight"]))
            w = torch.cat([torch.cat([hi, hi], dim=1), torch.cat([lo, torch.zeros_like(lo)], dim=1)], dim=0)
            T["conv_out.weight"] = torch.nn.functional.pad(w, (0, 0, 0, 16 - 2 * co)).contiguous()
        else:
            T["conv_out.weight"] = h(conv2d_w(P["conv_out.weight"], pad_out=16 - co))
        T["conv_out.bias"] = f(torch.nn.functional.pad(P["conv_out.bias"], (0, 16 - co)))
        return T, S, temb_total, xoff

    def _ensure_handle(self):
        if self._handle is not None:
            return
        L = _lib.lib()
        tensors, scalars, temb_total, xattn_total = self.pack_parameters()
        lay = block_layout(self._cfg)
        ints = [self._cfg["in_channels"], self._cfg["out_channels"], *lay["boc"], *lay["heads"],
                *[int(a) for a in lay["down_attn"]], self._cfg["layers_per_block"], self._cfg["cross_attention_dim"],
                self._cfg["addition_time_embed_dim"], temb_total, xattn_total]
        floats = [self.eps["cross"], self.eps["plain"], self.eps["mid"], self.eps["up"]]
        names = list(tensors)
        tn = (C.c_char_p * len(names))(*[n.encode() for n in names])
        tp = (C.c_void_p * len(names))(*[tensors[n].data_ptr() for n in names])
        sn_list = list(scalars)
        sn = (C.c_char_p * len(sn_list))(*[n.encode() for n in sn_list])
        sv = (C.c_double * len(sn_list))(*[scalars[n] for n in sn_list])
        handle = C.c_void_p()
        _lib.check(L.evw_unet_create(C.byref(handle), (C.c_int * len(ints))(*ints), len(ints),
                                     (C.c_float * len(floats))(*floats), len(floats), tn, tp, len(names), sn, sv,
                                     len(sn_list)), "evw_unet_create")
        self._handle, self._packed = handle, tensors

    def free_master_parameters(self):
        """Drop the fp32 diffusers-layout copy once packed (inference only needs the packed fp16 set)."""
        self._ensure_handle()
        self._params = {}

    def _workspace(self, B, T, h, w):
        key = (B, T, h, w)
        if key not in self._ws:
            n = _lib.lib().evw_unet_workspace_bytes(self._handle, B, T, h, w)
            if n < 0:
                _lib.check(-1, "evw_unet_workspace_bytes")
            # one live shape at a time.  The plan wants a 1024-byte aligned base (swizzled TMA tiles); the caching
            # allocator only promises 512, so over-allocate and hand out the first aligned offset (the base tensor
            # stays alive through the view).
            self._ws = {}
            base = torch.empty(n + 1024, dtype=torch.uint8, device=self._device)
            off = (-base.data_ptr()) % 1024
            self._ws = {key: base[off:off + n]}
        return self._ws[key]

    def plan_info(self) -> Tuple[int, float]:
        launches, flops = C.c_int64(), C.c_double()
        _lib.check(_lib.lib().evw_unet_plan_info(self._handle, C.byref(launches), C.byref(flops)), "evw_unet_plan_info")
        return launches.value, flops.value

    def graph_replays(self) -> int:
        """Calls served by replaying the captured CUDA graph of the current plan (-1 before the first call)."""
        return int(_lib.lib().evw_unet_graph_replays(self._handle)) if self._handle is not None else -1

    def gn_fused(self) -> int:
        """GroupNorms of the current plan whose statistics come from the producing GEMM's epilogue (-1 before the first call)."""
        return int(_lib.lib().evw_unet_gn_fused(self._handle)) if self._handle is not None else -1

    # ------------------------------------------------------------------ compute
    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int], encoder_hidden_states: torch.Tensor,
                added_time_ids: torch.Tensor, return_dict: bool = True):
        """sample (batch, num_frames, channel, height, width); timestep scalar; encoder_hidden_states
        (batch, 1, cross_attention_dim); added_time_ids (batch, 3)."""
        _lib.require_cuda(sample, "sample")
        if sample.dim() != 5 or sample.shape[2] != self._cfg["in_channels"]:
            raise ValueError(f"sample must be [B, T, {self._cfg['in_channels']}, h, w], got {tuple(sample.shape)}")
        if torch.device(sample.device) != self._device:
            raise RuntimeError("sample and UNet parameters live on different devices")
        self._ensure_handle()
        B, T_, _, h, w = sample.shape
        t = float(timestep.reshape(-1)[0].item()) if torch.is_tensor(timestep) else float(timestep)
        x = sample.float().contiguous()
        ehs = encoder_hidden_states.to(self._device, torch.float32).reshape(B, -1).contiguous()
        if ehs.shape[1] != self._cfg["cross_attention_dim"]:
            raise ValueError("encoder_hidden_states must be [B, 1, cross_attention_dim] (single image-embedding token)")
        ids = added_time_ids.to(self._device, torch.float32).contiguous()
        out = torch.empty((B, T_, self._cfg["out_channels"], h, w), dtype=torch.float32, device=self._device)
        ws = self._workspace(B, T_, h, w)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().evw_unet_forward(self._handle, _lib.ptr(x), t, _lib.ptr(ehs), _lib.ptr(ids), _lib.ptr(out),
                                                   B, T_, h, w, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self._device)),
                       "evw_unet_forward")
        out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNetSpatioTemporalConditionOutput(sample=out)

    __call__ = forward

    @torch.no_grad()
    def denoise_step(self, latents: torch.Tensor, cond_latents: torch.Tensor, sigma: float, sigma_next: float,
                     encoder_hidden_states: torch.Tensor, added_time_ids: torch.Tensor, min_guidance: float = 1.0,
                     max_guidance: float = 3.0) -> torch.Tensor:
        """One loop iteration of pipeline_evoworld.py:689-725, in place on `latents` fp32 [1,T,4,h,w];
        cond_latents fp32 [2,T,in_channels-4,h,w] (CFG batch: row 0 = unconditional)."""
        _lib.require_cuda(latents, "latents")
        if latents.dtype != torch.float32 or not latents.is_contiguous():
            raise ValueError("latents must be a contiguous fp32 tensor (updated in place)")
        self._ensure_handle()
        _, T_, cl, h, w = latents.shape
        if cl != 4 or tuple(cond_latents.shape) != (2, T_, self._cfg["in_channels"] - 4, h, w):
            raise ValueError("latents must be [1,T,4,h,w] and cond_latents [2,T,in_channels-4,h,w]")
        cond = cond_latents.float().contiguous()
        ehs = encoder_hidden_states.to(self._device, torch.float32).reshape(2, -1).contiguous()
        ids = added_time_ids.to(self._device, torch.float32).contiguous()
        ws = self._workspace(2, T_, h, w)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().evw_denoise_step(self._handle, _lib.ptr(latents), _lib.ptr(cond), float(sigma),
                                                   float(sigma_next), _lib.ptr(ehs), _lib.ptr(ids), float(min_guidance),
                                                   float(max_guidance), T_, h, w, _lib.ptr(ws), ws.numel(),
                                                   _lib.stream_ptr(self._device)), "evw_denoise_step")
        return latents


def algorithmic_flops(cfg: dict, B: int, T: int, h: int, w: int) -> Dict[str, float]:
    """2*MAC count of one UNet forward on a [B,T,*,h,w] batch, walking the architecture as the reference
    executes it (every conv / linear / SDPA, including the full single-key cross-attention projections);
    SURVEY Appendix C.  Returns TFLOP per bucket and the total."""
    lay = block_layout(cfg)
    boc = lay["boc"]
    temb, cross = boc[0] * 4, cfg["cross_attention_dim"]
    BF = B * T
    hs, ws = [h], [w]
    for _ in range(3):
        hs.append((hs[-1] + 1) // 2); ws.append((ws[-1] + 1) // 2)
    level_of = {}
    for i in range(4):
        level_of[f"down_blocks.{i}"] = i
        level_of[f"up_blocks.{i}"] = 3 - i
    level_of["mid_block"] = 3
    conv = gemm = sdpa_s = sdpa_t = 0.0
    conv += 2.0 * BF * hs[0] * ws[0] * 9 * cfg["in_channels"] * boc[0]
    gemm += 2.0 * B * (boc[0] * temb + temb * temb + cfg["projection_class_embeddings_input_dim"] * temb + temb * temb)
    for p, cin, cout in lay["res"]:
        l = level_of[p.rsplit(".resnets", 1)[0]]
        M = BF * hs[l] * ws[l]
        conv += 2.0 * M * 9 * (cin * cout + cout * cout) + (2.0 * M * cin * cout if cin != cout else 0.0)
        conv += 2.0 * M * 3 * 2 * cout * cout
        gemm += 2.0 * BF * temb * cout * 2
    for p, c, heads in lay["att"]:
        l = level_of[p.rsplit(".attentions", 1)[0]]
        S = hs[l] * ws[l]
        M = BF * S
        gemm += 2.0 * M * c * c * 2                      # proj_in, proj_out
        gemm += 2.0 * T * (c * 4 * c + 4 * c * c) * B    # time_pos_embed MLP
        for _blk in range(2):                            # spatial + temporal block
            gemm += 2.0 * M * c * c * 4                  # attn1 q,k,v,out
            gemm += 2.0 * M * c * c * 2                  # attn2 to_q, to_out on every token
        gemm += 2.0 * BF * cross * c * 2                 # spatial attn2 to_k, to_v: one context token per frame
        gemm += 2.0 * (M // T) * cross * c * 2           # temporal attn2 to_k, to_v: one context token per (b, s)
        gemm += 2.0 * M * (c * 8 * c + 4 * c * c) * 3    # ff, ff_in, ff
        sdpa_s += 4.0 * BF * heads * S * S * 64 + 4.0 * BF * heads * S * 1 * 64
        sdpa_t += 4.0 * B * S * heads * T * T * 64 + 4.0 * B * S * heads * T * 1 * 64
    for p, c in lay["samplers"]:
        i = int(p.split(".")[1])
        l = i + 1 if p.startswith("down") else 3 - i - 1
        conv += 2.0 * BF * hs[l] * ws[l] * 9 * c * c
    conv += 2.0 * BF * hs[0] * ws[0] * 9 * boc[0] * cfg["out_channels"]
    tot = conv + gemm + sdpa_s + sdpa_t
    return {"conv": conv / 1e12, "gemm": gemm / 1e12, "sdpa_spatial": sdpa_s / 1e12, "sdpa_temporal": sdpa_t / 1e12,
            "total": tot / 1e12}
