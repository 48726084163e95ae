"""EVW_VAE_DEBUG=1 python tools/vae_debug.py [encode|decode]: per-op output statistics of a small VAE call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoworld_b200 import vae as V
dev = torch.device("cuda:0")
m = V.AutoencoderKLTemporalDecoder(block_out_channels=(64, 128, 128, 128)).init_random(0, dev)
torch.manual_seed(0)
if len(sys.argv) < 2 or sys.argv[1] == "encode":
    x = torch.rand(1, 3, 64, 128, device=dev) * 2 - 1
    d = m.encode(x).latent_dist
    print("encode finite:", bool(torch.isfinite(d.parameters).all()))
else:
    z = torch.randn(2, 4, 8, 8, device=dev)
    y = m.decode(z, num_frames=2).sample
    print("decode finite:", bool(torch.isfinite(y).all()))
