"""Block-level comparison of the UNet plan against the oracle: EVW_UNET_DEBUG + EVW_UNET_DUMP_DIR dump
every op output; forward hooks on the oracle give the matching block outputs."""
import os, sys, math, shutil, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
dump = tempfile.mkdtemp()
os.environ["EVW_UNET_DEBUG"] = "1"
os.environ["EVW_UNET_DUMP_DIR"] = dump
import torch
from test_gpu_unet import make_pair, SMALL, rel_l2

dev = torch.device("cuda:0")
oracle, ours = make_pair(SMALL, dev)
torch.manual_seed(1)
B, T, h, w = 2, 3, 16, 32
x = torch.randn(B, T, 18, h, w, device=dev)
ehs = torch.randn(B, 1, 64, device=dev)
ids = torch.tensor([[6.0, 127.0, 0.02]] * B, device=dev)
t = 0.25 * math.log(3.7)
captured = {}
def hook(name):
    def f(mod, inp, out):
        o = out[0] if isinstance(out, tuple) else out
        captured[name] = o.detach().float().cpu()
    return f
for name, mod in oracle.named_modules():
    if name and (name == "conv_in" or name.endswith("spatial_res_block") or name.endswith("temporal_res_block") or ".resnets." in name and name.count(".") == 3
                 or ".attentions." in name and name.count(".") == 3 or "samplers.0" in name and name.count(".") == 3
                 or name.startswith("mid_block.resnets.") and name.count(".") == 2 or name.startswith("mid_block.attentions.") and name.count(".") == 2
                 or name.endswith("transformer_blocks.0") or name.endswith(".proj_in") or name.endswith("norm1") or name.endswith(".attn1") or name.endswith(".ff")):
        mod.register_forward_hook(hook(name))
with torch.no_grad():
    want = oracle(x, t, ehs, ids)
got = ours(x, t, ehs, ids).sample
print("final rel_l2", rel_l2(got, want))

def to_rows(o):  # oracle [BF,C,h,w] or [B,C,T,h,w] or [BF,S,C] -> [rows, C]
    if o.dim() == 4:
        return o.permute(0, 2, 3, 1).reshape(-1, o.shape[1])
    if o.dim() == 5:
        return o.permute(0, 2, 3, 4, 1).reshape(-1, o.shape[1])
    return o.reshape(-1, o.shape[-1])

# label of the op whose output equals an oracle module output
match = {}
for name in captured:
    if name == "conv_in": match["conv_in"] = name
    elif name.endswith("spatial_res_block"): match[name + ".conv2"] = name
    elif ".resnets." in name and not name.endswith("res_block"): match[name + ".temporal_res_block.conv2"] = name
    elif "samplers.0" in name: match[name + ".conv"] = name
    elif name.endswith(".proj_in"): match[name] = name
    elif ".attentions." in name and name.count(".") <= 3 and not name.endswith("proj_in"): match[name + ".proj_out"] = name
    elif name.endswith("temporal_transformer_blocks.0"): pass
    elif name.endswith("transformer_blocks.0"): match[name + ".ff.net.2"] = name
for line in open(os.path.join(dump, "index.txt")):
    parts = line.split()
    i, label, n, f16 = parts[0], parts[1], parts[-2], parts[-1]
    key = label
    if key in match:
        arr = np.fromfile(os.path.join(dump, f"op{i}." + ("f16" if f16 == "1" else "f32")), dtype=np.float16 if f16 == "1" else np.float32)
        ref = to_rows(captured[match[key]])
        a = torch.from_numpy(arr.astype(np.float32)).reshape(ref.shape[0], -1)[:, : ref.shape[1]]
        print(f"op {i:>4s} {key:75s} rel_l2 {rel_l2(a, ref):.3e}")
shutil.rmtree(dump)
