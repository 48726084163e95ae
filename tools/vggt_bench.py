"""VGGT-1B forward on the GPU: evoworld_b200.vggt vs PyTorch eager (the oracle restatement with SDPA, fp32 and under bf16
autocast as the reference runs it, unified_loop_consistency.py:131-136).  Random weights, synthetic frames at the loop's
392 x 518.  python tools/vggt_bench.py [--frames 25] [--steps 3] [--no-eager] [--out gpurun_out/vggt_bench.json]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from evoworld_b200 import ops  # noqa: E402
from evoworld_b200 import vggt as V  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return sum(ts) / len(ts), min(ts)


class Counter:
    """FLOPs (GEMMs: 2 M N K; attention: 4 S^2 d per head and sequence) and launches of one forward."""

    def __init__(self):
        self.flops = {"gemm": 0.0, "attention": 0.0}
        self.launches = 0
        self._orig = {}

    def __enter__(self):
        c = self
        names = ("gemm_f16", "spatial_attention", "small_attention", "layer_norm", "layer_norm_f32", "activation_f16", "relu_inplace_f16",
                 "qknorm_rope_", "bilinear_ac", "adaln_modulate", "dpt_activate")
        for n in names:
            self._orig[n] = getattr(ops, n)

        def wrap(n):
            f = self._orig[n]

            def g(*a, **k):
                c.launches += 1
                if n == "gemm_f16":
                    a0, w = a[0], a[1]
                    rows = a0.numel() // a0.shape[-1]
                    c.flops["gemm"] += 2.0 * rows * w.shape[0] * w.shape[1]
                elif n == "spatial_attention":
                    _, frames, S, heads = a
                    c.flops["attention"] += 4.0 * frames * heads * S * S * 64
                elif n == "small_attention":
                    _, B, S, heads, hd, _ = a
                    c.flops["attention"] += 4.0 * B * heads * S * S * hd
                return f(*a, **k)
            return g
        for n in names:
            setattr(ops, n, wrap(n))
        return self

    def __exit__(self, *e):
        for n, f in self._orig.items():
            setattr(ops, n, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--height", type=int, default=392)
    ap.add_argument("--width", type=int, default=518)
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-point-head", action="store_true")
    ap.add_argument("--profile-once", action="store_true", help="one warm forward, then one forward between cudaProfilerStart / Stop (ncu --profile-from-start off)")
    ap.add_argument("--dpt-chunk", type=int, default=32, help="frames per DPT-head pass (the reference: 8)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = dict(V.DEFAULT_CONFIG)
    if a.no_point_head:
        cfg["point_head"] = False
    sd = V.random_state_dict(cfg, seed=0, device=dev)
    m = V.VGGT(**cfg).to(dev)
    m.load_state_dict(sd)
    m._pack()
    images = torch.rand((1, a.frames, 3, a.height, a.width), device=dev)
    if a.profile_once:
        m.free_master_parameters()
        m(images)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        m(images)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    with Counter() as c:
        m(images)
    torch.cuda.synchronize()
    res = {"frames": a.frames, "size": [a.height, a.width], "params": m.num_parameters(), "launches": c.launches,
           "tflop": {k: v / 1e12 for k, v in c.flops.items()}}
    ms, best = timed(lambda: m(images, frames_chunk_size=a.dpt_chunk), a.steps, a.warmup)
    tf = sum(c.flops.values()) / 1e12
    res["native"] = {"ms": ms, "ms_best": best, "frames_per_s": a.frames / ms * 1e3, "tflops": tf / ms * 1e3}
    # sections
    last = cfg["depth"] - 1
    keep = set(cfg["dpt_layers"]) | {last}
    sec = {}
    sec["aggregator_ms"], _ = timed(lambda: m._aggregate(images, keep=keep), max(1, a.steps - 1), 1)
    pairs, dims = m._aggregate(images, keep=keep)
    B, S, P, H, W = dims
    sec["camera_head_ms"], _ = timed(lambda: m._camera(pairs[last], B, S, P), max(1, a.steps - 1), 1)
    sec["depth_head_ms"], _ = timed(lambda: m._dpt_chunked(pairs, dims, "depth_head.", "exp", 2, a.dpt_chunk), max(1, a.steps - 1), 1)
    res["sections"] = sec
    del pairs
    print(json.dumps(res), flush=True)
    if not a.no_eager:
        from oracle import vggt_torch as O

        O.FUSED_ATTN = True
        m.free_master_parameters()
        torch.cuda.empty_cache()
        ocfg = {k: v for k, v in cfg.items()}
        with torch.no_grad():
            for name, ctx in (("eager_bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16)), ("eager_fp32_tf32", None)):
                torch.backends.cuda.matmul.allow_tf32 = True
                torch.backends.cudnn.allow_tf32 = True

                def run():
                    if ctx is None:
                        return O.vggt_forward(images, sd, ocfg, point_head=cfg["point_head"])
                    with ctx:                                   # the reference: aggregator under autocast, heads in fp32 (models/vggt.py:65)
                        toks, start = O.aggregator(images, sd, ocfg)
                    toks = [t.float() for t in toks]
                    out = {"pose_enc": O.camera_head(toks[-1], sd, ocfg)[-1]}
                    out["depth"] = O.dpt_head(toks, a.height, a.width, start, sd, ocfg, "depth_head.", "exp")
                    if cfg["point_head"]:
                        out["wp"] = O.dpt_head(toks, a.height, a.width, start, sd, ocfg, "point_head.", "inv_log")
                    return out
                try:
                    ms_e, best_e = timed(run, max(1, a.steps - 1), 1)
                    res[name] = {"ms": ms_e, "ms_best": best_e, "speedup": ms_e / ms}
                except torch.OutOfMemoryError as e:  # noqa: PERF203
                    res[name] = {"error": "out of memory"}
                    torch.cuda.empty_cache()
    print(json.dumps(res), flush=True)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
