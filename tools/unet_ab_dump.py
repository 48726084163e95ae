"""A/B dump of every op of one UNet forward under two environments (e.g. EVW_GEMM_RES_TMA=0 vs 1): runs itself twice as a
subprocess with EVW_UNET_DEBUG + EVW_UNET_DUMP_DIR and reports the first ops whose outputs differ.
    python tools/unet_ab_dump.py "EVW_GEMM_RES_TMA=0" "EVW_GEMM_RES_TMA=1" [B T h w]"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("EVW_AB_CHILD"):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import math, torch
    from test_gpu_unet import make_pair, SMALL
    dev = torch.device("cuda:0")
    _, ours = make_pair(SMALL, dev)
    B, T, h, w = [int(v) for v in os.environ["EVW_AB_SHAPE"].split(",")]
    torch.manual_seed(1)
    x = torch.randn(B, T, 18, h, w, device=dev); ehs = torch.randn(B, 1, 64, device=dev)
    ids = torch.tensor([[6.0, 127.0, 0.02]] * B, device=dev)
    out = ours(x, 0.25 * math.log(3.7), ehs, ids).sample
    np.save(os.path.join(os.environ["EVW_UNET_DUMP_DIR"], "final.npy"), out.cpu().numpy())
    sys.exit(0)
envs = sys.argv[1:3]
shape = sys.argv[3:7] if len(sys.argv) >= 7 else ["2", "14", "24", "40"]
dirs = []
for e in envs:
    d = tempfile.mkdtemp(); dirs.append(d)
    env = dict(os.environ, EVW_AB_CHILD="1", EVW_UNET_DEBUG="1", EVW_UNET_DUMP_DIR=d, EVW_AB_SHAPE=",".join(shape), EVW_UNET_GRAPH="0")
    for kv in e.split():
        k, v = kv.split("="); env[k] = v
    subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=True, stderr=subprocess.DEVNULL)
idx = [l.split() for l in open(os.path.join(dirs[0], "index.txt"))]
shown = 0
for parts in idx:
    i, label, n, fp16 = parts[0], " ".join(parts[1:-2]), parts[-2], parts[-1]
    ext = ".f16" if fp16 == "1" else ".f32"
    a = np.fromfile(os.path.join(dirs[0], f"op{i}{ext}"), dtype=np.float16 if fp16 == "1" else np.float32).astype(np.float64)
    b = np.fromfile(os.path.join(dirs[1], f"op{i}{ext}"), dtype=np.float16 if fp16 == "1" else np.float32).astype(np.float64)
    d = np.linalg.norm(a - b) / (np.linalg.norm(a) + 1e-30)
    if d > 1e-7 and shown < 12:
        bad = np.nonzero(np.abs(a - b) > 1e-6 * (np.abs(a) + 1e-3))[0]
        print(f"op {i} {label} n={n}: rel diff {d:.3e}, {bad.size} elements differ, first at {bad[:6]}, a={a[bad[:3]]}, b={b[bad[:3]]}")
        shown += 1
fa, fb = np.load(os.path.join(dirs[0], "final.npy")), np.load(os.path.join(dirs[1], "final.npy"))
print("final rel diff", np.linalg.norm(fa - fb) / np.linalg.norm(fa))
