"""Per-op time profile of one denoise step (EVW_UNET_PROFILE): serialised CUDA-event timing of every planned op,
aggregated by op class.  Shares, not absolutes (ops are synchronised one by one)."""
import os, sys, re, collections, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 14
prof = os.path.join(ROOT, "gpurun_out", f"unet_profile_T{T}.txt")
import torch
import bench_denoise as bd
from evoworld_b200.unet import UNetSpatioTemporalConditionModel
dev = torch.device("cuda:0")
unet = UNetSpatioTemporalConditionModel(**bd.UNET_CFG).init_random(0, dev)
unet._ensure_handle(); unet.free_master_parameters()
lat, cond, ehs, ids = [t.to(dev) for t in bd.make_inputs(T, 72, 128, dev, 0)]
for _ in range(2):
    unet.denoise_step(lat.clone(), cond, 700.0, 545.7, ehs, ids)
torch.cuda.synchronize()
if os.path.exists(prof):
    os.remove(prof)
os.environ["EVW_UNET_PROFILE"] = prof
unet.denoise_step(lat.clone(), cond, 700.0, 545.7, ehs, ids)
torch.cuda.synchronize()
del os.environ["EVW_UNET_PROFILE"]
rows = []
for line in open(prof):
    if line.startswith("END"):
        break
    i, ms, label = line.split(" ", 2)
    rows.append((int(i), float(ms), label.strip()))
def cls(label):
    l = label.split(" ")[0]
    if "sdpa" in l: return "attention.temporal" if "temporal_transformer" in l else "attention.spatial"
    if re.search(r"norm(1|2|3|_in)?$", l) or l == "conv_norm_out": return "layernorm" if "transformer_blocks" in l else "groupnorm"
    if "temporal_res_block.conv" in l: return "conv.temporal"
    if "spatial_res_block.conv" in l or "samplers" in l or l in ("conv_in", "conv_out"): return "conv.3x3"
    if "ff.net.0" in l or "ff_in.net.0" in l: return "gemm.geglu"
    if "ff.net.2" in l or "ff_in.net.2" in l: return "gemm.ff2"
    if "qkv" in l: return "gemm.qkv"
    if "to_out" in l or "proj_in" in l or "proj_out" in l: return "gemm.proj"
    return "other:" + l
agg = collections.defaultdict(lambda: [0, 0.0])
for _, ms, label in rows:
    a = agg[cls(label)]; a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
print(f"T={T}: {len(rows)} ops, serialised total {tot:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} n={v[0]:4d} {v[1]:9.2f} ms  {100 * v[1] / tot:5.1f}%")
print("top ops:")
for i, ms, label in sorted(rows, key=lambda r: -r[1])[:25]:
    print(f"  {i:4d} {ms:8.3f} ms {label}")
# GEMM shapes: time against the FLOPs of the shape (labels end in "<rows>x<N>x<K_total>"), sorted by the time above a
# 1300 TFLOP/s floor and above the HBM floor of the operands that must move (A fp16 + output [+ fp32 residual])
shapes = collections.defaultdict(lambda: [0, 0.0, ""])
for _, ms, label in rows:
    m = re.search(r" (\d+)x(\d+)x(\d+)$", label)
    if not m:
        continue
    M, N, K = (int(g) for g in m.groups())
    key = (cls(label), M, N, K)
    s = shapes[key]; s[0] += 1; s[1] += ms; s[2] = label.split(" ")[0]
print("GEMM shapes (class, rows, N, K): calls, total ms, TFLOP/s, ms above the 1300 TFLOP/s floor")
out = []
for (c, M, N, K), (n, ms, ex) in shapes.items():
    fl = 2.0 * M * N * K * n
    floor = fl / 1300e12 * 1e3
    out.append((ms - floor, c, M, N, K, n, ms, fl / ms / 1e9, ex))
for lost, c, M, N, K, n, ms, tf, ex in sorted(out, reverse=True)[:40]:
    print(f"  {c:14s} {M:7d} x {N:5d} x {K:6d}  n={n:3d} {ms:8.3f} ms {tf:7.0f} TFLOP/s  +{lost:6.2f} ms   e.g. {ex}")
print(f"  sum of GEMM time above the floor: {sum(o[0] for o in out):.1f} ms of {sum(o[6] for o in out):.1f} ms")
