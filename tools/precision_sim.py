"""CPU precision ledger for the denoise step (no GPU needed).

Replays the fp32 oracle UNet (oracle/unet_torch.py) in float64 while emulating, layer category by layer
category, what the sm_100a path does to the operands: every GEMM / convolution operand (activation and
weight) rounded to fp16 ("f16") or split into an fp16 head + fp16 tail ("split": the tail carries the rounding
error of the head, so the product keeps ~22 significant bits), q/k/v/P/attention output stored as fp16,
conv1 outputs stored as fp16.  Prints the relative L2 of the output against the float64 oracle for
 (a) everything emulated (what the kernels do),
 (b) one category at a time emulated (its stand-alone contribution),
 (c) everything emulated but one category exact (what promoting that category buys).
Used to pick which (cheap) layers get the split-precision treatment; the GPU numbers are in profiles/.
"""
from __future__ import annotations

import argparse
import math
import sys
from contextlib import contextmanager
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import unet_torch as O  # noqa: E402

MODE = {}          # category -> "exact" | "f16" | "split"
DEFAULT = ["exact"]


def r16(x):
    return x.to(torch.float16).to(x.dtype)


def rsplit(x):
    hi = r16(x)
    return hi + r16(x - hi)


def q(x, cat):
    m = MODE.get(cat, DEFAULT[0])
    if m == "exact":
        return x
    if m == "split":
        return rsplit(x)
    return r16(x)


def category(name: str, mod: nn.Module) -> str:
    lvl = ""
    parts = name.split(".")
    if parts[0] == "down_blocks":
        lvl = f"L{parts[1]}d"
    elif parts[0] == "up_blocks":
        lvl = f"L{3 - int(parts[1])}u"
    elif parts[0] == "mid_block":
        lvl = "L3m"
    if name in ("conv_in", "conv_out"):
        return name
    if "embedding" in name or "time_emb_proj" in name or "time_pos_embed" in name:
        return "emb"
    if "attn2" in name:
        return "xattn"
    kind = None
    if "spatial_res_block.conv1" in name: kind = "sconv1"
    elif "spatial_res_block.conv2" in name: kind = "sconv2"
    elif "spatial_res_block.conv_shortcut" in name: kind = "sconv2"
    elif "temporal_res_block.conv1" in name: kind = "tconv1"
    elif "temporal_res_block.conv2" in name: kind = "tconv2"
    elif "samplers" in name: kind = "resample"
    elif name.endswith("proj_in"): kind = "proj_in"
    elif name.endswith("proj_out"): kind = "proj_out"
    elif "temporal_transformer_blocks" in name:
        if "attn1.to_out" in name: kind = "t_to_out"
        elif "attn1" in name: kind = "t_qkv"
        elif "net.0" in name: kind = "t_geglu"
        elif "net.2" in name: kind = "t_ff2"
    elif "transformer_blocks" in name:
        if "attn1.to_out" in name: kind = "s_to_out"
        elif "attn1" in name: kind = "s_qkv"
        elif "net.0" in name: kind = "s_geglu"
        elif "net.2" in name: kind = "s_ff2"
    if kind is None:
        raise KeyError(name)
    return f"{kind}@{lvl}"


def install(model: nn.Module):
    cats = {}
    for name, mod in model.named_modules():
        if isinstance(mod, (nn.Linear, nn.Conv2d, nn.Conv3d)):
            cat = category(name, mod)
            cats[name] = cat

            def fwd(x, mod=mod, cat=cat):
                w = q(mod.weight, cat)
                x = q(x, cat)
                if isinstance(mod, nn.Linear):
                    return F.linear(x, w, mod.bias)
                if isinstance(mod, nn.Conv2d):
                    return F.conv2d(x, w, mod.bias, mod.stride, mod.padding)
                return F.conv3d(x, w, mod.bias, mod.stride, mod.padding)

            mod.forward = fwd
        elif isinstance(mod, O.Attention) and "attn2" not in name:
            lvl = category(name + ".to_q", mod.to_q).split("@")[1]
            pre = "t" if "temporal" in name else "s"
            cat = f"{pre}_sdpa@{lvl}"
            cats[name] = cat

            def afwd(x, context=None, mod=mod, cat=cat):
                b, n, _ = x.shape
                qq = q(mod.to_q(x), cat).view(b, n, mod.heads, -1).transpose(1, 2)
                k = q(mod.to_k(x), cat).view(b, n, mod.heads, -1).transpose(1, 2)
                v = q(mod.to_v(x), cat).view(b, n, mod.heads, -1).transpose(1, 2)
                s = (qq @ k.transpose(-1, -2)) * (qq.shape[-1] ** -0.5)
                # the kernel rounds the un-normalised exp(s - m) to fp16, sums in fp32, divides at the end
                m = s.amax(-1, keepdim=True)
                p = torch.exp(s - m)
                l = p.sum(-1, keepdim=True)
                o = (q(p, cat) @ v) / l
                o = q(o.transpose(1, 2).reshape(b, n, -1), cat)
                return mod.to_out[0](o)

            mod.forward = afwd
    # conv1 outputs (+temb) are stored as fp16 (h16) before the second GroupNorm
    for name, mod in model.named_modules():
        if isinstance(mod, O.ResnetBlock2D):
            lvl = category(name + ".conv1", mod.conv1).split("@")[1]
            cat = f"h16s@{lvl}"
            cats[name + ".h16"] = cat

            def rfwd(x, temb, mod=mod, cat=cat):
                h = mod.conv1(F.silu(mod.norm1(x)))
                h = q(h + mod.time_emb_proj(F.silu(temb))[:, :, None, None], cat)
                h = mod.conv2(F.silu(mod.norm2(h)))
                if mod.conv_shortcut is not None:
                    x = mod.conv_shortcut(x)
                return x + h

            mod.forward = rfwd
        elif isinstance(mod, O.TemporalResnetBlock):
            lvl = category(name + ".conv1", mod.conv1).split("@")[1]
            cat = f"h16t@{lvl}"
            cats[name + ".h16"] = cat

            def tfwd(x, temb, mod=mod, cat=cat):
                h = mod.conv1(F.silu(mod.norm1(x)))
                t = mod.time_emb_proj(F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
                h = q(h + t, cat)
                h = mod.conv2(F.silu(mod.norm2(h)))
                return x + h

            mod.forward = tfwd
    return cats


def rel(a, b):
    return float((a - b).norm() / b.norm())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-width", action="store_true")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--T", type=int, default=3)
    ap.add_argument("--h", type=int, default=16)
    ap.add_argument("--w", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--group", default="kind", choices=["kind", "kind_level", "level"])
    ap.add_argument("--promote", default="", help="comma list of category prefixes forced to --promote-mode in the 'all' run")
    ap.add_argument("--promote-mode", default="split")
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    if args.full_width:
        cfg = dict(in_channels=18, block_out_channels=(320, 640, 1280, 1280), num_attention_heads=(5, 10, 20, 20), cross_attention_dim=1024)
    else:
        cfg = dict(in_channels=18, block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), cross_attention_dim=64)
    model = O.UNetSpatioTemporalConditionModel(**cfg).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
            if n.endswith("mix_factor"):
                p.copy_(torch.randn_like(p))
    model = model.double()
    cats = install(model)
    torch.manual_seed(1)
    x = torch.randn(args.B, args.T, 18, args.h, args.w, dtype=torch.float64)
    ehs = torch.randn(args.B, 1, cfg["cross_attention_dim"], dtype=torch.float64)
    ids = torch.tensor([[6.0, 127.0, 0.02]] * args.B, dtype=torch.float64)
    t = 0.25 * math.log(3.7)

    def run():
        with torch.no_grad():
            return model(x, t, ehs, ids)

    def key(c):
        if args.group == "kind":
            return c.split("@")[0]
        if args.group == "level":
            return c.split("@")[1] if "@" in c else c
        return c

    MODE.clear(); DEFAULT[0] = "exact"
    want = run()
    all_cats = sorted(set(cats.values()))
    groups = sorted(set(key(c) for c in all_cats))
    DEFAULT[0] = "f16"
    base = rel(run(), want)
    print(f"all f16 (kernel emulation): {base:.4e}")
    if args.promote:
        pre = tuple(args.promote.split(","))
        for c in all_cats:
            if c.startswith(pre):
                MODE[c] = args.promote_mode
        e = rel(run(), want)
        print(f"with {args.promote} -> {args.promote_mode}: {e:.4e}")
        return
    rows = []
    for g in groups:
        MODE.clear(); DEFAULT[0] = "exact"
        for c in all_cats:
            if key(c) == g:
                MODE[c] = "f16"
        alone = rel(run(), want)
        MODE.clear(); DEFAULT[0] = "f16"
        for c in all_cats:
            if key(c) == g:
                MODE[c] = "exact"
        without = rel(run(), want)
        rows.append((g, alone, without))
        print(f"{g:16s} alone {alone:.3e}   all-but {without:.3e}  (gain {base - without:+.2e})", flush=True)
    print("quadrature sum of stand-alone contributions:", f"{math.sqrt(sum(a * a for _, a, _ in rows)):.4e}")


if __name__ == "__main__":
    main()
