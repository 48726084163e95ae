"""CLIP ViT image encoder on sm_100a.

Mirrors `transformers.CLIPVisionModelWithProjection` as the reference pipeline holds it (`self.image_encoder`,
evoworld/pipeline/pipeline_evoworld.py:238,262,289): `from_pretrained(path)` of a transformers-format folder
(config.json + model.safetensors / pytorch_model.bin, the library's state-dict keys), `.config`, `parameters()`, and
`__call__(pixel_values [B,3,H,W]) -> output.image_embeds [B, projection_dim]`.  Arithmetic (transformers
models/clip/modeling_clip.py: CLIPVisionEmbeddings, CLIPEncoderLayer, CLIPAttention, CLIPMLP): patch embedding, q/k/v (one
fused GEMM), out_proj, fc1, fc2 and the projection on the tcgen05 GEMM with fp16 operands / fp32 accumulation and the
residual adds in its epilogue, LayerNorms and the attention over the 257 tokens on their own kernels; the residual
stream stays fp32.  There is no CPU fallback.
"""
from __future__ import annotations

import json
import math
import os
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib, ops

DEFAULT_CONFIG = dict(  # laion CLIP ViT-H/14, the image encoder of Stable Video Diffusion
    hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, num_channels=3, image_size=224,
    patch_size=14, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5,
)


def param_spec(cfg) -> "OrderedDict[str, tuple]":
    d, i, p = cfg["hidden_size"], cfg["intermediate_size"], cfg["patch_size"]
    n_pos = (cfg["image_size"] // p) ** 2 + 1
    s: "OrderedDict[str, tuple]" = OrderedDict()
    e = "vision_model.embeddings."
    s[e + "class_embedding"] = (d,)
    s[e + "patch_embedding.weight"] = (d, cfg["num_channels"], p, p)
    s[e + "position_embedding.weight"] = (n_pos, d)
    s["vision_model.pre_layrnorm.weight"] = (d,); s["vision_model.pre_layrnorm.bias"] = (d,)
    for l in range(cfg["num_hidden_layers"]):
        b = f"vision_model.encoder.layers.{l}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[b + f"self_attn.{n}.weight"] = (d, d); s[b + f"self_attn.{n}.bias"] = (d,)
        s[b + "layer_norm1.weight"] = (d,); s[b + "layer_norm1.bias"] = (d,)
        s[b + "mlp.fc1.weight"] = (i, d); s[b + "mlp.fc1.bias"] = (i,)
        s[b + "mlp.fc2.weight"] = (d, i); s[b + "mlp.fc2.bias"] = (d,)
        s[b + "layer_norm2.weight"] = (d,); s[b + "layer_norm2.bias"] = (d,)
    s["vision_model.post_layernorm.weight"] = (d,); s["vision_model.post_layernorm.bias"] = (d,)
    s["visual_projection.weight"] = (cfg["projection_dim"], d)
    return s


class CLIPVisionModelWithProjection:
    config_name = "config.json"

    def __init__(self, **kwargs):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        if cfg["hidden_size"] % 64 or cfg["intermediate_size"] % 64 or cfg["hidden_size"] % cfg["num_attention_heads"]:
            raise NotImplementedError("hidden_size and intermediate_size must be multiples of 64 (tcgen05 GEMM K blocks)")
        if cfg["hidden_act"] not in ("gelu", "quick_gelu"):
            raise NotImplementedError(f"hidden_act={cfg['hidden_act']!r}")
        if cfg["image_size"] % cfg["patch_size"]:
            raise ValueError("image_size must be a multiple of patch_size")
        self._cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self._spec = param_spec(cfg)
        self._params: Dict[str, torch.Tensor] = {}
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._device = torch.device("cpu")

    # ------------------------------------------------------------------ parameters
    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return torch.float32

    def num_parameters(self) -> int:
        return sum(math.prod(s) for s in self._spec.values())

    def parameters(self):
        return iter(self._params.values())

    def named_parameters(self):
        return iter(self._params.items())

    def state_dict(self):
        return OrderedDict((k, self._params[k]) for k in self._spec)

    def load_state_dict(self, sd, strict: bool = True):
        ignore = ("vision_model.embeddings.position_ids",)  # a buffer older transformers versions save
        missing = [k for k in self._spec if k not in sd]
        unexpected = [k for k in sd if k not in self._spec and k not in ignore]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shape in self._spec.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {shape}")
                self._params[k] = sd[k].detach().to(self._device, torch.float32, copy=True).contiguous()
        self._packed = None
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def init_random(self, seed: int = 0, device=None):
        dev = torch.device(device) if device is not None else self._device
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        self._params = {}
        for name, shape in self._spec.items():
            if "norm" in name:
                t = torch.ones(shape, device=dev) if name.endswith("weight") else torch.zeros(shape, device=dev)
            else:
                t = torch.randn(shape, device=dev, generator=g) * 0.02
            self._params[name] = t
        self._device, self._packed = dev, None
        return self

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
            self._packed = None
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    @classmethod
    def from_config(cls, config: dict):
        config = config.get("vision_config", config)  # a full CLIPConfig nests the vision tower's
        return cls(**{k: v for k, v in config.items() if k in DEFAULT_CONFIG})

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, **kwargs):
        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else str(pretrained_model_name_or_path)
        cfg_path = os.path.join(root, cls.config_name)
        if not os.path.isfile(cfg_path):
            raise FileNotFoundError(f"{cfg_path} not found")
        with open(cfg_path) as f:
            model = cls.from_config(json.load(f))
        for stem in ("model.safetensors", "model.fp16.safetensors", "pytorch_model.bin"):
            path = os.path.join(root, stem)
            if os.path.isfile(path):
                if stem.endswith(".safetensors"):
                    from safetensors.torch import load_file

                    sd = load_file(path)
                else:
                    sd = torch.load(path, map_location="cpu", weights_only=True)
                model.load_state_dict(sd)
                return model
        raise FileNotFoundError(f"no weights found under {root}")

    def save_pretrained(self, save_directory: str):
        from safetensors.torch import save_file

        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump({**self._cfg, "architectures": ["CLIPVisionModelWithProjection"], "model_type": "clip_vision_model"}, f, indent=2)
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}, os.path.join(save_directory, "model.safetensors"))

    # ------------------------------------------------------------------ packing
    @torch.no_grad()
    def _pack(self):
        if self._packed is not None:
            return self._packed
        if self._device.type != "cuda":
            raise RuntimeError("evoworld_b200 CLIP: parameters must be on a CUDA device (no CPU fallback)")
        if len(self._params) != len(self._spec):
            raise RuntimeError("CLIP parameters are not initialised (load_state_dict / init_random first)")
        P, cfg = self._params, self._cfg
        h = lambda t: t.to(torch.float16).contiguous()
        f = lambda t: t.to(torch.float32).contiguous()
        T: Dict[str, torch.Tensor] = {}
        e = "vision_model.embeddings."
        w = P[e + "patch_embedding.weight"].reshape(cfg["hidden_size"], -1)          # [d, 3 p p] in (c, ky, kx) order = unfold's
        self._k_patch = w.shape[1]
        kp = (w.shape[1] + 63) // 64 * 64                                             # K padded to whole 64-wide GEMM blocks
        T["patch.weight"] = h(torch.nn.functional.pad(w, (0, kp - w.shape[1])))
        T["class_pos"] = f(P[e + "position_embedding.weight"])                        # [n_pos, d]; row 0 gets the class embedding added
        T["class_pos"][0] += P[e + "class_embedding"]
        for n in ("pre_layrnorm", "post_layernorm"):
            T[n + ".weight"] = f(P[f"vision_model.{n}.weight"]); T[n + ".bias"] = f(P[f"vision_model.{n}.bias"])
        for l in range(cfg["num_hidden_layers"]):
            b = f"vision_model.encoder.layers.{l}."
            T[f"{l}.qkv.weight"] = h(torch.cat([P[b + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")], 0))
            T[f"{l}.qkv.bias"] = f(torch.cat([P[b + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")], 0))
            for src, dst in (("self_attn.out_proj", "out"), ("mlp.fc1", "fc1"), ("mlp.fc2", "fc2")):
                T[f"{l}.{dst}.weight"] = h(P[b + src + ".weight"]); T[f"{l}.{dst}.bias"] = f(P[b + src + ".bias"])
            for n in ("layer_norm1", "layer_norm2"):
                T[f"{l}.{n}.weight"] = f(P[b + n + ".weight"]); T[f"{l}.{n}.bias"] = f(P[b + n + ".bias"])
        pw = P["visual_projection.weight"]
        self._proj_rows = pw.shape[0]
        T["proj.weight"] = h(torch.nn.functional.pad(pw, (0, 0, 0, (-pw.shape[0]) % 8)))  # GEMM N is a multiple of 8
        self._packed = T
        return T

    def free_master_parameters(self):
        self._pack()
        self._params = {}

    # ------------------------------------------------------------------ compute
    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, output_hidden_states: bool = False, return_dict: bool = True):
        """pixel_values [B,3,H,W] (CLIP-normalised, H = W = image_size) -> .image_embeds [B, projection_dim] (and
        .last_hidden_state [B, 1 + patches, hidden])."""
        _lib.require_cuda(pixel_values, "pixel_values")
        cfg = self._cfg
        B, Cc, H, W = pixel_values.shape
        if Cc != cfg["num_channels"] or H != cfg["image_size"] or W != cfg["image_size"]:
            raise ValueError(f"pixel_values must be [B,{cfg['num_channels']},{cfg['image_size']},{cfg['image_size']}], got {tuple(pixel_values.shape)}")
        T = self._pack()
        d, heads, p = cfg["hidden_size"], cfg["num_attention_heads"], cfg["patch_size"]
        hd = d // heads
        eps = float(cfg["layer_norm_eps"])
        n_patch = (H // p) * (W // p)
        S = n_patch + 1
        # CLIPVisionEmbeddings: Conv2d(k = stride = patch, no bias) as one GEMM over the unfolded patches
        x = pixel_values.to(torch.float32)
        cols = torch.nn.functional.unfold(x, kernel_size=p, stride=p).transpose(1, 2).reshape(B * n_patch, -1)   # [B n, 3 p p]
        a = torch.zeros((B * n_patch, T["patch.weight"].shape[1]), dtype=torch.float16, device=x.device)
        a[:, : cols.shape[1]] = cols
        patches = ops.gemm_f16(a, T["patch.weight"], out_dtype=torch.float32).view(B, n_patch, d)
        hs = torch.zeros((B, S, d), dtype=torch.float32, device=x.device)
        hs[:, 1:] = patches
        hs = (hs + T["class_pos"][None]).reshape(B * S, d).contiguous()     # class token row = class_embedding + pos[0]
        hs = ops.layer_norm_f32(hs, T["pre_layrnorm.weight"], T["pre_layrnorm.bias"], eps)   # the fp32 residual stream
        scale = hd ** -0.5
        for l in range(cfg["num_hidden_layers"]):
            n1 = ops.layer_norm(hs, T[f"{l}.layer_norm1.weight"], T[f"{l}.layer_norm1.bias"], eps)
            qkv = ops.gemm_f16(n1, T[f"{l}.qkv.weight"], bias=T[f"{l}.qkv.bias"], out_dtype=torch.float16)
            att = ops.small_attention(qkv, B, S, heads, hd, scale)
            hs = ops.gemm_f16(att, T[f"{l}.out.weight"], bias=T[f"{l}.out.bias"], res1=hs, out_dtype=torch.float32)
            n2 = ops.layer_norm(hs, T[f"{l}.layer_norm2.weight"], T[f"{l}.layer_norm2.bias"], eps)
            if cfg["hidden_act"] == "gelu":   # GELU in the GEMM epilogue: no fp32 intermediate, no activation pass
                act = ops.gemm_f16(n2, T[f"{l}.fc1.weight"], bias=T[f"{l}.fc1.bias"], act="gelu", out_dtype=torch.float16)
            else:
                f1 = ops.gemm_f16(n2, T[f"{l}.fc1.weight"], bias=T[f"{l}.fc1.bias"], out_dtype=torch.float32)
                act = ops.activation_f16(f1, cfg["hidden_act"])
            hs = ops.gemm_f16(act, T[f"{l}.fc2.weight"], bias=T[f"{l}.fc2.bias"], res1=hs, out_dtype=torch.float32)
        last = hs.view(B, S, d)
        pooled = ops.layer_norm(last[:, 0].contiguous(), T["post_layernorm.weight"], T["post_layernorm.bias"], eps)   # fp16 [B, d]
        embeds = ops.gemm_f16(pooled, T["proj.weight"], out_dtype=torch.float32)[:, : self._proj_rows].contiguous()
        out = SimpleNamespace(image_embeds=embeds.to(pixel_values.dtype), last_hidden_state=last)
        return out if return_dict else (out.image_embeds, out.last_hidden_state)

    __call__ = forward
