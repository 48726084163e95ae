"""VGGT forward pass on sm_100a behind the reference's `VGGT` surface.

The caller (unified_loop_consistency.py:114-136) does `VGGT().to(device).eval()`, `load_state_dict(model.pt)` and
`model(images [1,S,3,H,W])` and reads `pose_enc`, `depth`, `depth_conf` (+ `world_points`, `world_points_conf`, `images`)
from the returned dict (:352-366) — this class offers exactly that, with the reference module's own state-dict keys
(third_party/vggt/vggt/models/vggt.py:17-25; `track_head.*` keys are accepted and ignored: the track head only runs with
`query_points`, which the reference loop never passes).

Arithmetic (file:line under third_party/vggt/vggt):
  * aggregator  models/aggregator.py:184-306 — DINOv2 ViT patch tokens (layers/vision_transformer.py:215-275), camera /
    register tokens, `depth` x (frame block, global block) with q/k LayerNorm + 2-D RoPE (layers/attention.py:50-72,
    layers/rope.py:116-188), LayerScale (folded into the proj / fc2 weights at pack time)
  * camera head heads/camera_head.py:75-140 — 4 refinement iterations of adaLN-modulated trunk blocks
  * DPT heads   heads/dpt_head.py:159-272 — depth (exp) and point (inv_log) heads
Every linear and convolution runs on the tcgen05 implicit GEMM (fp16 operands, fp32 accumulation, bias / LayerScale /
residual / positional-embedding adds in its epilogue), frame and global attention (head width 64; global = one sequence of
S x P tokens) on the tcgen05 flash-attention kernel, the rest on the small kernels of csrc/vggt_elem.cu, csrc/clip_elem.cu
and the LayerNorm of csrc/unet_elem.cu.  The residual stream stays fp32.  torch is used for memory, views / permutes /
concatenations of activations and one-off parameter preparation (weight packing, the bicubic resize of the DINOv2
position table, sin / cos tables).  There is no CPU fallback.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib, ops

RESNET_MEAN = (0.485, 0.456, 0.406)
RESNET_STD = (0.229, 0.224, 0.225)

DEFAULT_CONFIG = dict(  # facebook/VGGT-1B: models/vggt.py:17-25 with every constructor default
    img_size=518, patch_size=14, embed_dim=1024, depth=24, num_heads=16, num_register_tokens=4, rope_freq=100.0,
    vit_depth=24, vit_heads=16,                      # dinov2_vitl14_reg (layers/vision_transformer.py:363-374)
    camera_heads=16, camera_trunk_depth=4, camera_iterations=4,
    dpt_features=256, dpt_out_channels=(256, 512, 1024, 1024), dpt_layers=(4, 11, 17, 23), point_head=True,
)


def _block_spec(s, pre, d, hidden, qk_norm):
    s[pre + "norm1.weight"] = (d,); s[pre + "norm1.bias"] = (d,)
    s[pre + "attn.qkv.weight"] = (3 * d, d); s[pre + "attn.qkv.bias"] = (3 * d,)
    if qk_norm:
        for n in ("q_norm", "k_norm"):
            s[pre + f"attn.{n}.weight"] = (64,); s[pre + f"attn.{n}.bias"] = (64,)
    s[pre + "attn.proj.weight"] = (d, d); s[pre + "attn.proj.bias"] = (d,)
    s[pre + "ls1.gamma"] = (d,)
    s[pre + "norm2.weight"] = (d,); s[pre + "norm2.bias"] = (d,)
    s[pre + "mlp.fc1.weight"] = (hidden, d); s[pre + "mlp.fc1.bias"] = (hidden,)
    s[pre + "mlp.fc2.weight"] = (d, hidden); s[pre + "mlp.fc2.bias"] = (d,)
    s[pre + "ls2.gamma"] = (d,)


def param_spec(cfg) -> "OrderedDict[str, tuple]":
    """Names and shapes of the reference VGGT.state_dict() (without track_head), in the reference's order."""
    d, p, r = cfg["embed_dim"], cfg["patch_size"], cfg["num_register_tokens"]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    a = "aggregator."
    s[a + "camera_token"] = (1, 2, 1, d)
    s[a + "register_token"] = (1, 2, r, d)
    v = a + "patch_embed."
    s[v + "cls_token"] = (1, 1, d)
    s[v + "pos_embed"] = (1, (cfg["img_size"] // p) ** 2 + 1, d)
    s[v + "register_tokens"] = (1, r, d)
    s[v + "mask_token"] = (1, d)
    s[v + "patch_embed.proj.weight"] = (d, 3, p, p); s[v + "patch_embed.proj.bias"] = (d,)
    for i in range(cfg["vit_depth"]):
        _block_spec(s, f"{v}blocks.{i}.", d, 4 * d, False)
    s[v + "norm.weight"] = (d,); s[v + "norm.bias"] = (d,)
    for kind in ("frame_blocks", "global_blocks"):
        for i in range(cfg["depth"]):
            _block_spec(s, f"{a}{kind}.{i}.", d, 4 * d, True)
    c, dc = "camera_head.", 2 * d
    s[c + "empty_pose_tokens"] = (1, 1, 9)
    for i in range(cfg["camera_trunk_depth"]):
        _block_spec(s, f"{c}trunk.{i}.", dc, 4 * dc, False)
    for n in ("token_norm", "trunk_norm"):
        s[c + n + ".weight"] = (dc,); s[c + n + ".bias"] = (dc,)
    s[c + "embed_pose.weight"] = (dc, 9); s[c + "embed_pose.bias"] = (dc,)
    s[c + "poseLN_modulation.1.weight"] = (3 * dc, dc); s[c + "poseLN_modulation.1.bias"] = (3 * dc,)
    s[c + "pose_branch.fc1.weight"] = (dc // 2, dc); s[c + "pose_branch.fc1.bias"] = (dc // 2,)
    s[c + "pose_branch.fc2.weight"] = (9, dc // 2); s[c + "pose_branch.fc2.bias"] = (9,)
    heads = [("point_head.", 4)] if cfg["point_head"] else []
    heads.append(("depth_head.", 2))
    f, oc = cfg["dpt_features"], tuple(cfg["dpt_out_channels"])
    for h, out_dim in heads:
        s[h + "norm.weight"] = (dc,); s[h + "norm.bias"] = (dc,)
        for j in range(4):
            s[f"{h}projects.{j}.weight"] = (oc[j], dc, 1, 1); s[f"{h}projects.{j}.bias"] = (oc[j],)
        s[h + "resize_layers.0.weight"] = (oc[0], oc[0], 4, 4); s[h + "resize_layers.0.bias"] = (oc[0],)
        s[h + "resize_layers.1.weight"] = (oc[1], oc[1], 2, 2); s[h + "resize_layers.1.bias"] = (oc[1],)
        s[h + "resize_layers.3.weight"] = (oc[3], oc[3], 3, 3); s[h + "resize_layers.3.bias"] = (oc[3],)
        for j in range(4):
            s[f"{h}scratch.layer{j + 1}_rn.weight"] = (f, oc[j], 3, 3)
        for j in (1, 2, 3, 4):
            rn = f"{h}scratch.refinenet{j}."
            s[rn + "out_conv.weight"] = (f, f, 1, 1); s[rn + "out_conv.bias"] = (f,)
            for u in (("resConfUnit1", "resConfUnit2") if j != 4 else ("resConfUnit2",)):
                for cv in ("conv1", "conv2"):
                    s[f"{rn}{u}.{cv}.weight"] = (f, f, 3, 3); s[f"{rn}{u}.{cv}.bias"] = (f,)
        s[h + "scratch.output_conv1.weight"] = (f // 2, f, 3, 3); s[h + "scratch.output_conv1.bias"] = (f // 2,)
        s[h + "scratch.output_conv2.0.weight"] = (32, f // 2, 3, 3); s[h + "scratch.output_conv2.0.bias"] = (32,)
        s[h + "scratch.output_conv2.2.weight"] = (out_dim, 32, 1, 1); s[h + "scratch.output_conv2.2.bias"] = (out_dim,)
    return s


def random_state_dict(cfg, seed: int = 0, device="cpu") -> "OrderedDict[str, torch.Tensor]":
    """Seeded fp32 parameters (on the CPU: deterministic for a given torch build — the golden vectors use these): fan-in
    scaled weights, small biases, norm weights around 1, LayerScale gammas around 0.3 so that every block moves the
    residual stream."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in param_spec(cfg).items():
        r = torch.randn(shape, generator=g, device=device)
        leaf = name.rsplit(".", 1)[-1]
        if "norm" in name and leaf == "weight":
            t = 1.0 + 0.1 * r
        elif leaf == "gamma":
            t = 0.3 + 0.1 * r
        elif leaf == "bias":
            t = 0.05 * r
        elif leaf == "weight":
            if name.endswith("resize_layers.0.weight") or name.endswith("resize_layers.1.weight"):
                fan_in = shape[0]                       # ConvTranspose2d [in, out, k, k] with stride = k: one tap per output
            else:
                fan_in = math.prod(shape[1:])
            t = r / math.sqrt(fan_in)
        elif leaf == "pos_embed":
            t = 0.2 * r
        else:                                            # camera / register / cls / mask / empty-pose tokens
            t = 0.5 * r
        sd[name] = t
    return sd


def _rope_tables(max_pos: int, freq: float, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """layers/rope.py:84-114 for 32 features per direction: fp32 [max_pos, 16] cos / sin of pos * freq^(-2j/32)."""
    exponents = torch.arange(0, 32, 2, device=device).float() / 32
    inv_freq = 1.0 / (freq ** exponents)
    positions = torch.arange(max_pos, device=device, dtype=inv_freq.dtype)
    angles = torch.einsum("i,j->ij", positions, inv_freq)
    return angles.cos().contiguous(), angles.sin().contiguous()


def _uv_pos_embed(w: int, h: int, C: int, aspect: float, device, ratio: float = 0.1) -> torch.Tensor:
    """heads/dpt_head.py:274-283 + heads/utils.py:11-109: [h * w, C] fp32 = ratio x sinusoidal embedding of the uv grid."""
    diag = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / diag, 1.0 / diag
    xs = torch.linspace(-sx * (w - 1) / w, sx * (w - 1) / w, steps=w, dtype=torch.float32, device=device)
    ys = torch.linspace(-sy * (h - 1) / h, sy * (h - 1) / h, steps=h, dtype=torch.float32, device=device)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")

    def sincos(D, pos):
        omega = torch.arange(D // 2, dtype=torch.double, device=device)
        omega /= D / 2.0
        omega = 1.0 / 100 ** omega
        out = torch.einsum("m,d->md", pos.reshape(-1).double(), omega)
        return torch.cat([torch.sin(out), torch.cos(out)], dim=1).float()

    emb = torch.cat([sincos(C // 2, uu), sincos(C // 2, vv)], dim=-1)
    return (emb * ratio).reshape(h * w, C).contiguous()


def _conv_w(w: torch.Tensor, n_pad: int = 0, c_pad: int = 0) -> torch.Tensor:
    """Conv2d weight [N, C, kh, kw] -> tap-major fp16 [N (+pad), kh kw (C + pad)] (ops.CONV3x3_TAPS order: ky outer, kx inner)."""
    N, C, kh, kw = w.shape
    w = w.permute(0, 2, 3, 1)                                   # [N, kh, kw, C]
    if c_pad:
        w = F.pad(w, (0, c_pad))
    w = w.reshape(N, -1)
    if n_pad:
        w = F.pad(w, (0, 0, 0, n_pad))
    return w.to(torch.float16).contiguous()


def _check_device(dev: torch.device):
    if dev.type != "cuda":
        raise RuntimeError("evoworld_b200 VGGT: parameters must be on a CUDA device (no CPU fallback)")


class VGGT:
    def __init__(self, img_size: int = 518, patch_size: int = 14, embed_dim: int = 1024, **overrides):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim)
        unknown = [k for k in overrides if k not in cfg]
        if unknown:
            raise TypeError(f"VGGT: unknown configuration keys {unknown}")
        cfg.update(overrides)
        d = cfg["embed_dim"]
        if d != 64 * cfg["num_heads"] or d != 64 * cfg["vit_heads"]:
            raise NotImplementedError("VGGT: embed_dim must be 64 x num_heads (= 64 x vit_heads): the tcgen05 attention kernel is built for head width 64")
        if (2 * d) % cfg["camera_heads"] or 2 * d // cfg["camera_heads"] > 256:
            raise NotImplementedError("VGGT: camera head width must divide 2 x embed_dim and be <= 256")
        f = cfg["dpt_features"]
        if f % 128 or any(c % 64 for c in cfg["dpt_out_channels"]) or len(cfg["dpt_out_channels"]) != 4 or len(cfg["dpt_layers"]) != 4:
            raise NotImplementedError("VGGT: dpt_features must be a multiple of 128 and the four dpt_out_channels multiples of 64")
        if max(cfg["dpt_layers"]) >= cfg["depth"]:
            raise ValueError("VGGT: dpt_layers beyond the aggregator depth")
        self._cfg = cfg
        self._spec = param_spec(cfg)
        self._params: Dict[str, torch.Tensor] = {}
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._cache: Dict[tuple, torch.Tensor] = {}
        self._device = torch.device("cpu")

    # ------------------------------------------------------------------ parameters (nn.Module-like surface)
    @property
    def config(self):
        return dict(self._cfg)

    @property
    def device(self):
        return self._device

    def num_parameters(self) -> int:
        return sum(math.prod(s) for s in self._spec.values())

    def parameters(self):
        return iter(self._params.values())

    def named_parameters(self):
        return iter(self._params.items())

    def state_dict(self):
        return OrderedDict((k, self._params[k]) for k in self._spec if k in self._params)

    def load_state_dict(self, sd, strict: bool = True):
        missing = [k for k in self._spec if k not in sd]
        unexpected = [k for k in sd if k not in self._spec and not k.startswith("track_head.")]
        if not self._cfg["point_head"]:
            unexpected = [k for k in unexpected if not k.startswith("point_head.")]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shape in self._spec.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {shape}")
                self._params[k] = sd[k].detach().to(self._device, torch.float32, copy=True).contiguous()
        self._packed = None
        self._cache = {}
        from types import SimpleNamespace

        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def init_random(self, seed: int = 0, device=None):
        if device is not None:
            self._device = torch.device(device)
        self.load_state_dict(random_state_dict(self._cfg, seed))
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
            self._cache = {}
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
        """A local folder holding model.safetensors / model.pt (the files of facebook/VGGT-1B) or such a file itself."""
        root = str(pretrained_model_name_or_path)
        cands = [root] if os.path.isfile(root) else [os.path.join(root, n) for n in ("model.safetensors", "model.pt")]
        for path in cands:
            if os.path.isfile(path):
                if path.endswith(".safetensors"):
                    from safetensors.torch import load_file

                    sd = load_file(path)
                else:
                    sd = torch.load(path, map_location="cpu", weights_only=True)
                model = cls(**kwargs)
                model.load_state_dict(sd)
                return model
        raise FileNotFoundError(f"no VGGT weights (model.safetensors / model.pt) under {root}; there is no network download")

    # ------------------------------------------------------------------ packing
    def _pack_block(self, T, pre):
        P = self._params
        h = lambda t: t.to(torch.float16).contiguous()
        f = lambda t: t.to(torch.float32).contiguous()
        for n in ("norm1", "norm2"):
            T[pre + n + ".weight"] = f(P[pre + n + ".weight"]); T[pre + n + ".bias"] = f(P[pre + n + ".bias"])
        T[pre + "qkv.weight"] = h(P[pre + "attn.qkv.weight"]); T[pre + "qkv.bias"] = f(P[pre + "attn.qkv.bias"])
        for n in ("q_norm", "k_norm"):
            if pre + f"attn.{n}.weight" in P:
                T[pre + n + ".weight"] = f(P[pre + f"attn.{n}.weight"]); T[pre + n + ".bias"] = f(P[pre + f"attn.{n}.bias"])
        g1, g2 = P[pre + "ls1.gamma"], P[pre + "ls2.gamma"]            # LayerScale folded: x + g * (W a + b) = x + (g W) a + g b
        T[pre + "proj.weight"] = h(g1[:, None] * P[pre + "attn.proj.weight"]); T[pre + "proj.bias"] = f(g1 * P[pre + "attn.proj.bias"])
        T[pre + "fc1.weight"] = h(P[pre + "mlp.fc1.weight"]); T[pre + "fc1.bias"] = f(P[pre + "mlp.fc1.bias"])
        T[pre + "fc2.weight"] = h(g2[:, None] * P[pre + "mlp.fc2.weight"]); T[pre + "fc2.bias"] = f(g2 * P[pre + "mlp.fc2.bias"])

    @torch.no_grad()
    def _pack(self):
        if self._packed is not None:
            return self._packed
        _check_device(self._device)
        if len(self._params) != len(self._spec):
            raise RuntimeError("VGGT parameters are not initialised (load_state_dict / init_random first)")
        P, cfg = self._params, self._cfg
        h = lambda t: t.to(torch.float16).contiguous()
        f = lambda t: t.to(torch.float32).contiguous()
        T: Dict[str, torch.Tensor] = {}
        v = "aggregator.patch_embed."
        w = P[v + "patch_embed.proj.weight"].reshape(cfg["embed_dim"], -1)             # (c, ky, kx) order = unfold's
        T["patch.weight"] = h(F.pad(w, (0, (-w.shape[1]) % 64)))
        T["patch.bias"] = f(P[v + "patch_embed.proj.bias"])
        T["vit.cls"] = f(P[v + "cls_token"][0, 0]); T["vit.reg"] = f(P[v + "register_tokens"][0]); T["vit.pos"] = f(P[v + "pos_embed"])
        T["agg.camera"] = f(P["aggregator.camera_token"][0, :, 0]); T["agg.register"] = f(P["aggregator.register_token"][0])
        for i in range(cfg["vit_depth"]):
            self._pack_block(T, f"{v}blocks.{i}.")
        T[v + "norm.weight"] = f(P[v + "norm.weight"]); T[v + "norm.bias"] = f(P[v + "norm.bias"])
        for kind in ("frame_blocks", "global_blocks"):
            for i in range(cfg["depth"]):
                self._pack_block(T, f"aggregator.{kind}.{i}.")
        c = "camera_head."
        for i in range(cfg["camera_trunk_depth"]):
            self._pack_block(T, f"{c}trunk.{i}.")
        for n in ("token_norm", "trunk_norm"):
            T[c + n + ".weight"] = f(P[c + n + ".weight"]); T[c + n + ".bias"] = f(P[c + n + ".bias"])
        dc = 2 * cfg["embed_dim"]
        T[c + "empty_pose"] = f(P[c + "empty_pose_tokens"].reshape(1, 9))
        T[c + "ones"] = torch.ones(dc, device=self._device); T[c + "zeros"] = torch.zeros(dc, device=self._device)
        T[c + "embed_pose.weight"] = h(F.pad(P[c + "embed_pose.weight"], (0, 64 - 9))); T[c + "embed_pose.bias"] = f(P[c + "embed_pose.bias"])
        T[c + "mod.weight"] = h(P[c + "poseLN_modulation.1.weight"]); T[c + "mod.bias"] = f(P[c + "poseLN_modulation.1.bias"])
        T[c + "pb1.weight"] = h(P[c + "pose_branch.fc1.weight"]); T[c + "pb1.bias"] = f(P[c + "pose_branch.fc1.bias"])
        T[c + "pb2.weight"] = h(F.pad(P[c + "pose_branch.fc2.weight"], (0, 0, 0, 16 - 9))); T[c + "pb2.bias"] = f(F.pad(P[c + "pose_branch.fc2.bias"], (0, 16 - 9)))
        for hd in (("point_head.", "depth_head.") if cfg["point_head"] else ("depth_head.",)):
            T[hd + "norm.weight"] = f(P[hd + "norm.weight"]); T[hd + "norm.bias"] = f(P[hd + "norm.bias"])
            for j in range(4):
                T[f"{hd}projects.{j}.weight"] = _conv_w(P[f"{hd}projects.{j}.weight"]); T[f"{hd}projects.{j}.bias"] = f(P[f"{hd}projects.{j}.bias"])
            for j, k in ((0, 4), (1, 2)):   # ConvTranspose2d(k, stride k) [in, out, k, k] -> GEMM rows (ky, kx, out), then a pixel shuffle
                wt = P[f"{hd}resize_layers.{j}.weight"]
                T[f"{hd}resize.{j}.weight"] = h(wt.permute(2, 3, 1, 0).reshape(k * k * wt.shape[1], wt.shape[0]))
                T[f"{hd}resize.{j}.bias"] = f(P[f"{hd}resize_layers.{j}.bias"].repeat(k * k))
            T[hd + "resize.3.weight"] = _conv_w(P[hd + "resize_layers.3.weight"]); T[hd + "resize.3.bias"] = f(P[hd + "resize_layers.3.bias"])
            for j in range(4):
                T[f"{hd}rn.{j}.weight"] = _conv_w(P[f"{hd}scratch.layer{j + 1}_rn.weight"])
            for j in (1, 2, 3, 4):
                rn = f"{hd}scratch.refinenet{j}."
                T[rn + "out_conv.weight"] = _conv_w(P[rn + "out_conv.weight"]); T[rn + "out_conv.bias"] = f(P[rn + "out_conv.bias"])
                for u in (("resConfUnit1", "resConfUnit2") if j != 4 else ("resConfUnit2",)):
                    for cv in ("conv1", "conv2"):
                        T[f"{rn}{u}.{cv}.weight"] = _conv_w(P[f"{rn}{u}.{cv}.weight"]); T[f"{rn}{u}.{cv}.bias"] = f(P[f"{rn}{u}.{cv}.bias"])
            sc = hd + "scratch."
            T[sc + "oc1.weight"] = _conv_w(P[sc + "output_conv1.weight"]); T[sc + "oc1.bias"] = f(P[sc + "output_conv1.bias"])
            T[sc + "oc2a.weight"] = _conv_w(P[sc + "output_conv2.0.weight"], n_pad=32)            # 32 -> 64 outputs (zero rows): the next K block
            T[sc + "oc2a.bias"] = f(F.pad(P[sc + "output_conv2.0.bias"], (0, 32)))
            w2 = P[sc + "output_conv2.2.weight"]
            T[sc + "oc2b.weight"] = _conv_w(w2, n_pad=16 - w2.shape[0], c_pad=32)
            T[sc + "oc2b.bias"] = f(F.pad(P[sc + "output_conv2.2.bias"], (0, 16 - w2.shape[0])))
        self._packed = T
        return T

    def free_master_parameters(self):
        self._pack()
        self._params = {}

    # ------------------------------------------------------------------ building blocks
    def _block(self, x: torch.Tensor, pre: str, eps: float, attn) -> torch.Tensor:
        """layers/block.py:84-107 (eval) on the fp32 residual stream x [rows, d]; attn(qkv fp16 [rows, 3d]) -> fp16 [rows, d]."""
        T = self._packed
        n1 = ops.layer_norm(x, T[pre + "norm1.weight"], T[pre + "norm1.bias"], eps)
        qkv = ops.gemm_f16(n1, T[pre + "qkv.weight"], bias=T[pre + "qkv.bias"], out_dtype=torch.float16)
        x = ops.gemm_f16(attn(qkv), T[pre + "proj.weight"], bias=T[pre + "proj.bias"], res1=x, out_dtype=torch.float32)
        n2 = ops.layer_norm(x, T[pre + "norm2.weight"], T[pre + "norm2.bias"], eps)
        f1 = ops.gemm_f16(n2, T[pre + "fc1.weight"], bias=T[pre + "fc1.bias"], act="gelu", out_dtype=torch.float16)   # GELU in the epilogue
        return ops.gemm_f16(f1, T[pre + "fc2.weight"], bias=T[pre + "fc2.bias"], res1=x, out_dtype=torch.float32)

    def _pos_tokens(self, h0: int, w0: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """DINOv2 position table resized to the h0 x w0 patch grid (vision_transformer.py:183-213: bicubic, antialias,
        interpolate_offset 0): (class-token row [d], patch rows [h0 w0, d]) — parameter preparation, cached per grid."""
        key = ("pos", h0, w0)
        if key not in self._cache:
            pe = self._packed["vit.pos"]
            N = pe.shape[1] - 1
            M = int(math.sqrt(N))
            patch = pe[:, 1:]
            if not (h0 * w0 == N and h0 == w0):
                patch = F.interpolate(patch.reshape(1, M, M, -1).permute(0, 3, 1, 2), mode="bicubic", antialias=True, size=(h0, w0))
                patch = patch.permute(0, 2, 3, 1).reshape(1, h0 * w0, -1)
            self._cache[key] = (pe[0, 0].contiguous(), patch[0].contiguous())
        return self._cache[key]

    # ------------------------------------------------------------------ aggregator
    @torch.no_grad()
    def _aggregate(self, images: torch.Tensor, keep: Optional[set] = None):
        """models/aggregator.py:184-258 -> {layer: (frame tokens, global tokens)} fp32 [B*S*P, d] each (kept layers only)."""
        cfg, T = self._cfg, self._pack()
        B, S, Cin, H, W = images.shape
        p, d, r = cfg["patch_size"], cfg["embed_dim"], cfg["num_register_tokens"]
        if Cin != 3:
            raise ValueError(f"Expected 3 input channels, got {Cin}")
        if H % p or W % p:
            raise ValueError(f"image size {H}x{W} must be a multiple of the patch size {p}")
        dev = images.device
        Fr, h0, w0 = B * S, H // p, W // p
        n = h0 * w0
        P = n + 1 + r
        # image normalisation (aggregator.py:201) + the 14 x 14 patches as rows of the patch-embedding GEMM, in one kernel
        a = ops.patchify_f16(images.reshape(Fr, 3, H, W).to(torch.float32).contiguous(), p, T["patch.weight"].shape[1], RESNET_MEAN, RESNET_STD)
        cls_pos, patch_pos = self._pos_tokens(h0, w0)
        v = "aggregator.patch_embed."
        # PatchEmbed conv (k = stride = patch) + bias + the patch rows of the position table in one GEMM epilogue
        patches = ops.gemm_f16(a, T["patch.weight"], bias=T["patch.bias"], rowvec=patch_pos, rv_div=1, rv_mod=n, out_dtype=torch.float32)
        tok = torch.empty((Fr, P, d), dtype=torch.float32, device=dev)
        tok[:, 0] = T["vit.cls"] + cls_pos
        tok[:, 1: 1 + r] = T["vit.reg"]
        tok[:, 1 + r:] = patches.view(Fr, n, d)
        xs = tok.view(Fr * P, d)
        heads = cfg["vit_heads"]
        for i in range(cfg["vit_depth"]):
            xs = self._block(xs, f"{v}blocks.{i}.", 1e-6, lambda qkv: ops.spatial_attention(qkv, Fr, P, heads))
        xs = ops.layer_norm_f32(xs, T[v + "norm.weight"], T[v + "norm.bias"], 1e-6).view(Fr, P, d)
        # camera / register tokens: index 0 for the first frame of every sequence, index 1 for the others (aggregator.py:309-331)
        first = torch.zeros(S, dtype=torch.long, device=dev)
        first[1:] = 1
        first = first.repeat(B)
        tok = torch.empty((Fr, P, d), dtype=torch.float32, device=dev)
        tok[:, 0] = T["agg.camera"][first]
        tok[:, 1: 1 + r] = T["agg.register"][first]
        tok[:, 1 + r:] = xs[:, 1 + r:]
        xs = tok.view(Fr * P, d)
        key = ("rope", h0, w0)
        if key not in self._cache:
            grid = torch.cartesian_prod(torch.arange(h0, device=dev), torch.arange(w0, device=dev)) + 1
            pos = torch.cat([torch.zeros(1 + r, 2, dtype=grid.dtype, device=dev), grid], 0).to(torch.int32).contiguous()
            self._cache[key] = (pos,) + _rope_tables(max(h0, w0) + 1, float(cfg["rope_freq"]), dev)
        pos, cos_t, sin_t = self._cache[key]
        heads = cfg["num_heads"]

        def attn(pre, frames, seq):
            def run(qkv):
                ops.qknorm_rope_(qkv, heads, P, pos, T[pre + "q_norm.weight"], T[pre + "q_norm.bias"], T[pre + "k_norm.weight"],
                                 T[pre + "k_norm.bias"], cos_t, sin_t, 1e-5)
                return ops.spatial_attention(qkv, frames, seq, heads)
            return run

        out = {}
        for i in range(cfg["depth"]):
            pre = f"aggregator.frame_blocks.{i}."
            xs = self._block(xs, pre, 1e-5, attn(pre, Fr, P))
            fr = xs
            pre = f"aggregator.global_blocks.{i}."
            xs = self._block(xs, pre, 1e-5, attn(pre, B, S * P))
            if keep is None or i in keep:
                out[i] = (fr, xs)
        return out, (B, S, P, H, W)

    @torch.no_grad()
    def aggregator(self, images: torch.Tensor) -> Tuple[List[torch.Tensor], int]:
        """The reference's `model.aggregator(images)`: (list of `depth` tensors fp32 [B, S, P, 2 embed_dim], patch_start_idx)."""
        _lib.require_cuda(images, "images")
        if images.dim() == 4:
            images = images[None]
        with torch.autocast(device_type=images.device.type, enabled=False):
            pairs, (B, S, P, _, _) = self._aggregate(images)
        d = self._cfg["embed_dim"]
        return [torch.cat([pairs[i][0].view(B, S, P, d), pairs[i][1].view(B, S, P, d)], dim=-1) for i in range(self._cfg["depth"])], \
            1 + self._cfg["num_register_tokens"]

    # ------------------------------------------------------------------ camera head
    @torch.no_grad()
    def _camera(self, pair, B: int, S: int, P: int) -> List[torch.Tensor]:
        """heads/camera_head.py:75-140 on the camera tokens of the last aggregator layer -> pose encodings fp32 [B, S, 9]."""
        cfg, T = self._cfg, self._packed
        d = cfg["embed_dim"]
        dc, heads = 2 * d, cfg["camera_heads"]
        hd = dc // heads
        c = "camera_head."
        rows = B * S
        cam = torch.cat([pair[0].view(rows, P, d)[:, 0], pair[1].view(rows, P, d)[:, 0]], dim=-1).contiguous()
        x = ops.layer_norm_f32(cam, T[c + "token_norm.weight"], T[c + "token_norm.bias"], 1e-5)
        xn = ops.layer_norm_f32(x, T[c + "ones"], T[c + "zeros"], 1e-6)       # adaln_norm: no affine (the same in every iteration)
        pred = None
        outs = []
        attn = lambda qkv: ops.small_attention(qkv, B, S, heads, hd, hd ** -0.5)
        for _ in range(cfg["camera_iterations"]):
            inp = T[c + "empty_pose"].expand(rows, 9) if pred is None else pred
            a = torch.zeros((rows, 64), dtype=torch.float16, device=x.device)
            a[:, :9] = inp
            emb = ops.gemm_f16(a, T[c + "embed_pose.weight"], bias=T[c + "embed_pose.bias"], out_dtype=torch.float32)
            mod = ops.gemm_f16(ops.activation_f16(emb, "silu"), T[c + "mod.weight"], bias=T[c + "mod.bias"], out_dtype=torch.float32)
            t = ops.adaln_modulate(xn, mod, x)
            for i in range(cfg["camera_trunk_depth"]):
                t = self._block(t, f"{c}trunk.{i}.", 1e-5, attn)
            tn = ops.layer_norm(t, T[c + "trunk_norm.weight"], T[c + "trunk_norm.bias"], 1e-5)
            h1 = ops.gemm_f16(tn, T[c + "pb1.weight"], bias=T[c + "pb1.bias"], act="gelu", out_dtype=torch.float16)
            delta = ops.gemm_f16(h1, T[c + "pb2.weight"], bias=T[c + "pb2.bias"], out_dtype=torch.float32)[:, :9]
            pred = delta.contiguous() if pred is None else pred + delta
            outs.append(torch.cat([pred[:, :7], torch.relu(pred[:, 7:])], dim=-1).view(B, S, 9))   # head_act.py:11-34: FoV through relu
        return outs

    # ------------------------------------------------------------------ DPT head
    def _conv3(self, a: torch.Tensor, w, bias=None, **kw) -> torch.Tensor:
        """a fp16 [F, h, w, C] -> [F*h*w, N]: 3x3 convolution, padding 1, on the implicit GEMM."""
        Fr, h, wd, C = a.shape
        return ops.gemm_f16(a.view(Fr, 1, h, wd, C), w, taps=ops.CONV3x3_TAPS, bias=bias, **kw)

    def _rcu(self, x: torch.Tensor, shape, pre: str, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ResidualConvUnit (dpt_head.py:375-410) on x fp32 [rows, C]: conv2(relu(conv1(relu(x)))) + relu(x) (+ extra).  The
        unit's activation is nn.ReLU(inplace=True) (:333), so `self.activation(x)` (:397) overwrites x and the skip connection
        (:410) adds relu(x); x is overwritten here the same way (no caller reads it afterwards)."""
        T = self._packed
        Fr, h, w, C = shape
        c1 = self._conv3(ops.relu_inplace_f16(x).view(Fr, h, w, C), T[pre + "conv1.weight"], T[pre + "conv1.bias"], out_dtype=torch.float32)
        return self._conv3(ops.activation_f16(c1, "relu").view(Fr, h, w, C), T[pre + "conv2.weight"], T[pre + "conv2.bias"], res1=x,
                           res2=extra, out_dtype=torch.float32)

    def _fusion(self, pre: str, x0: torch.Tensor, shape, x1: Optional[torch.Tensor], size) -> torch.Tensor:
        """FeatureFusionBlock (dpt_head.py:413-460): x0 (+ rcu1(x1)) -> rcu2 -> bilinear (align_corners) -> 1x1 out_conv."""
        T = self._packed
        Fr, h, w, C = shape
        out = x0 if x1 is None else self._rcu(x1, shape, pre + "resConfUnit1.", extra=x0)
        out = self._rcu(out, shape, pre + "resConfUnit2.")
        up = ops.bilinear_ac(out.view(Fr, h, w, C), size[0], size[1], torch.float16)
        return ops.gemm_f16(up.view(-1, C), T[pre + "out_conv.weight"], bias=T[pre + "out_conv.bias"], out_dtype=torch.float32)

    @torch.no_grad()
    def _dpt(self, pairs, dims, hd: str, activation: str, out_dim: int, f0: int, f1: int):
        """heads/dpt_head.py:159-272 for frames f0:f1 of every sequence -> (pts fp32 [B, n, H, W, out_dim-1], conf [B, n, H, W])."""
        cfg, T = self._cfg, self._packed
        B, S, P, H, W = dims
        p, d, r = cfg["patch_size"], cfg["embed_dim"], cfg["num_register_tokens"]
        ph, pw = H // p, W // p
        n = ph * pw
        nf = f1 - f0
        Fr = B * nf
        feat = cfg["dpt_features"]
        dev = pairs[cfg["dpt_layers"][0]][0].device
        aspect = W / H

        def pe(w_, h_, C):
            key = ("uv", w_, h_, C, W, H)
            if key not in self._cache:
                self._cache[key] = _uv_pos_embed(w_, h_, C, aspect, dev)
            return self._cache[key]

        feats = []
        for j, li in enumerate(cfg["dpt_layers"]):
            fr, gl = pairs[li]
            x = torch.cat([fr.view(B, S, P, d)[:, f0:f1, 1 + r:], gl.view(B, S, P, d)[:, f0:f1, 1 + r:]], dim=-1).reshape(Fr * n, 2 * d)
            xn = ops.layer_norm(x, T[hd + "norm.weight"], T[hd + "norm.bias"], 1e-5)
            oc = T[f"{hd}projects.{j}.weight"].shape[0]
            y = ops.gemm_f16(xn, T[f"{hd}projects.{j}.weight"], bias=T[f"{hd}projects.{j}.bias"], rowvec=pe(pw, ph, oc), rv_div=1, rv_mod=n,
                             out_dtype=torch.float16)                                     # 1x1 projection + bias + uv embedding
            if j < 2:                                                                     # ConvTranspose2d(k, stride k) = GEMM + pixel shuffle
                k = 4 if j == 0 else 2
                z = ops.gemm_f16(y, T[f"{hd}resize.{j}.weight"], bias=T[f"{hd}resize.{j}.bias"], out_dtype=torch.float16)
                y = z.view(Fr, ph, pw, k, k, oc).permute(0, 1, 3, 2, 4, 5).reshape(Fr, ph * k, pw * k, oc).contiguous()
            elif j == 2:
                y = y.view(Fr, ph, pw, oc)
            else:                                                                         # 3x3 stride-2 conv = every other pixel of the stride-1 conv
                z = self._conv3(y.view(Fr, ph, pw, oc), T[hd + "resize.3.weight"], T[hd + "resize.3.bias"], out_dtype=torch.float16)
                y = z.view(Fr, ph, pw, oc)[:, ::2, ::2].contiguous()
            feats.append(y)
        shapes = [(Fr, f.shape[1], f.shape[2], feat) for f in feats]
        ls = [self._conv3(f, T[f"{hd}rn.{j}.weight"], out_dtype=torch.float32) for j, f in enumerate(feats)]
        sc = hd + "scratch."
        out = self._fusion(sc + "refinenet4.", ls[3], shapes[3], None, shapes[2][1:3])
        out = self._fusion(sc + "refinenet3.", out, shapes[2], ls[2], shapes[1][1:3])
        out = self._fusion(sc + "refinenet2.", out, shapes[1], ls[1], shapes[0][1:3])
        h1, w1 = shapes[0][1:3]
        out = self._fusion(sc + "refinenet1.", out, shapes[0], ls[0], (2 * h1, 2 * w1))
        o1 = self._conv3(ops.activation_f16(out, "identity").view(Fr, 2 * h1, 2 * w1, feat), T[sc + "oc1.weight"], T[sc + "oc1.bias"],
                         out_dtype=torch.float32)
        Ho, Wo = ph * p, pw * p
        up = ops.bilinear_ac(o1.view(Fr, 2 * h1, 2 * w1, feat // 2), Ho, Wo, torch.float16, addend=pe(Wo, Ho, feat // 2))
        o2 = self._conv3(up, T[sc + "oc2a.weight"], T[sc + "oc2a.bias"], out_dtype=torch.float32)
        o3 = ops.gemm_f16(ops.activation_f16(o2, "relu"), T[sc + "oc2b.weight"], bias=T[sc + "oc2b.bias"], out_dtype=torch.float32)
        pts, conf = ops.dpt_activate(o3, out_dim, activation)
        return pts.view(B, nf, Ho, Wo, out_dim - 1), conf.view(B, nf, Ho, Wo)

    def _dpt_chunked(self, pairs, dims, hd, activation, out_dim, chunk):
        S = dims[1]
        if not chunk or chunk >= S:
            return self._dpt(pairs, dims, hd, activation, out_dim, 0, S)
        parts = [self._dpt(pairs, dims, hd, activation, out_dim, s0, min(S, s0 + chunk)) for s0 in range(0, S, chunk)]
        return torch.cat([q[0] for q in parts], dim=1), torch.cat([q[1] for q in parts], dim=1)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, images: torch.Tensor, query_points=None, frames_chunk_size: Optional[int] = 32) -> Dict[str, torch.Tensor]:
        """models/vggt.py:27-92: images [S,3,H,W] or [B,S,3,H,W] in [0,1] -> pose_enc [B,S,9], depth [B,S,H,W,1], depth_conf
        [B,S,H,W], world_points [B,S,H,W,3], world_points_conf [B,S,H,W], images.  frames_chunk_size: the DPT heads process
        this many frames at a time (heads/dpt_head.py:119-157 chunks by 8 to bound memory; frames are independent in the heads, so
        the chunk size does not change the result — 32 keeps a segment's 25 frames in one pass, whose low-resolution layers then
        fill the GPU; None = all at once)."""
        _lib.require_cuda(images, "images")
        if query_points is not None:
            raise NotImplementedError("VGGT track head (query_points) is not part of the reference loop and is not built")
        if images.dim() == 4:
            images = images[None]
        cfg = self._cfg
        last = cfg["depth"] - 1
        # the reference wraps the call in bf16 autocast (unified_loop_consistency.py:131-136); the precision here is fixed by the
        # kernels, and the few torch helpers (einsum of the sin / cos tables, interpolate of the position table) must stay fp32
        with torch.autocast(device_type=images.device.type, enabled=False):
            pairs, dims = self._aggregate(images, keep=set(cfg["dpt_layers"]) | {last})
            B, S, P, H, W = dims
            out = {"pose_enc": self._camera(pairs[last], B, S, P)[-1]}
            depth, out["depth_conf"] = self._dpt_chunked(pairs, dims, "depth_head.", "exp", 2, frames_chunk_size)
            out["depth"] = depth
            if cfg["point_head"]:
                out["world_points"], out["world_points_conf"] = self._dpt_chunked(pairs, dims, "point_head.", "inv_log", 4, frames_chunk_size)
        out["images"] = images
        return out

    __call__ = forward


@torch.no_grad()
def run_vggt_inference(model: VGGT, perspective_frames, lift_dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """`UnifiedLoopConsistencyPipeline.run_vggt_inference` (unified_loop_consistency.py:336-367) with everything resident on
    the device: RGB frames uint8 [N, H, W, 3] (a CUDA tensor, or a list of host arrays as the reference passes) -> the
    preprocessing of load_and_preprocess_images without the PNG round trip (image_ops.vggt_preprocess_u8, Pillow-exact) -> VGGT ->
    pose encoding to extrinsic / intrinsic (:352) -> batch dimension squeezed (:357-362) -> `world_points_from_depth` by the
    depth lift (:366, evw_lift_depth).  Returns CUDA tensors (the reference converts to numpy at this point)."""
    import numpy as np

    from . import image_ops
    from .geometry import pose_encoding_to_extri_intri
    from .lift import lift_depth_device

    if not isinstance(perspective_frames, torch.Tensor):
        perspective_frames = torch.from_numpy(np.stack([np.asarray(f).astype(np.uint8) for f in perspective_frames]))
    frames = perspective_frames.to(model.device, non_blocking=True)
    images = image_ops.vggt_preprocess_u8(frames)[None]
    preds = model(images)
    extr, intr = pose_encoding_to_extri_intri(preds["pose_enc"], images.shape[-2:])
    preds["extrinsic"], preds["intrinsic"] = extr, intr
    out = {k: (v[0] if isinstance(v, torch.Tensor) else v) for k, v in preds.items()}
    out["world_points_from_depth"] = lift_depth_device(out["depth"], out["extrinsic"], out["intrinsic"], out_dtype=lift_dtype)
    return out
