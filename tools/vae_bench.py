"""VAE at the BASELINE clip size (576x1024): encode of 1 + 25 frames and temporal decode of 25 frames in chunks of 8
(decode_chunk_size=8 as forward_evoworld.py:197 / navigator_evoworld.py:205), CUDA-event timed after a warm-up call,
beside the PyTorch restatement (oracle/vae_torch.py) on the same GPU in fp16 autocast and in fp32 (cuDNN / cuBLAS)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from evoworld_b200 import vae as V

dev = torch.device("cuda:0")
H, W = 576, 1024
m = V.AutoencoderKLTemporalDecoder().init_random(0, dev)
sd = m.state_dict()
torch.manual_seed(0)
x = torch.rand(26, 3, H, W, device=dev) * 2 - 1
z = torch.randn(25, 4, H // 8, W // 8, device=dev)
res = {}


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def decode_chunks(model_decode):
    outs = []
    for i in range(0, 25, 8):
        zi = z[i:i + 8]
        outs.append(model_decode(zi, zi.shape[0]))
    return outs


ms_enc = timed(lambda: m.encode(x))
l0, f0, g0 = m.plan_info(0)
ms_dec = timed(lambda: decode_chunks(lambda zi, n: m.decode(zi, num_frames=n).sample))
l1, f1, g1 = m.plan_info(1)
# plan_info describes the LAST plan (encode: 2 frames of the 26 = 8+8+8+2; decode: 1 frame of 8+8+8+1): scale per frame
enc_tflop = f0 / 2 * 26 / 1e12
print(f"encode 26 x 576x1024: {ms_enc:8.1f} ms   ~{enc_tflop:.1f} TFLOP  {enc_tflop / ms_enc * 1e3:7.1f} TFLOP/s   ({l0} launches in the last 2-frame plan, {g0} GroupNorms fed by epilogues)")
m.decode(z[:8], num_frames=8)
l8, f8, g8 = m.plan_info(1)
dec_tflop = (3 * f8 + f1) / 1e12
print(f"decode 25 x 576x1024 (chunks of 8): {ms_dec:8.1f} ms   {dec_tflop:.1f} TFLOP  {dec_tflop / ms_dec * 1e3:7.1f} TFLOP/s   ({l8} launches per 8-frame chunk, {g8} GroupNorms fed by epilogues)")
res.update(encode_ms=ms_enc, encode_tflop=enc_tflop, decode_ms=ms_dec, decode_tflop=dec_tflop, launches_decode_chunk=l8)
peak = torch.cuda.max_memory_allocated() / 2**30
print(f"peak device memory {peak:.1f} GiB")
res["peak_gib"] = peak

if "--no-eager" not in sys.argv:
    from oracle import vae_torch as O
    del m
    torch.cuda.empty_cache()
    with torch.device(dev):
        o = O.AutoencoderKLTemporalDecoder().eval()
    o.load_state_dict(sd)
    for name, ctx, tf32 in (("fp16 autocast", lambda: torch.autocast("cuda", dtype=torch.float16), False),
                            ("fp32 (TF32 on)", lambda: torch.autocast("cuda", enabled=False), True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        try:
            with torch.no_grad(), ctx():
                ms_e = timed(lambda: [o.encode_moments(x[i:i + 2]) for i in range(0, 26, 2)], reps=1)
                ms_d = timed(lambda: decode_chunks(lambda zi, n: o.decode(zi, n)), reps=1)
            print(f"PyTorch eager {name:16s}: encode {ms_e:8.1f} ms  decode {ms_d:8.1f} ms   ours {ms_e / ms_enc:.2f}x / {ms_d / ms_dec:.2f}x faster")
            res["eager " + name] = dict(encode_ms=ms_e, decode_ms=ms_d)
        except torch.OutOfMemoryError as exc:
            print(f"PyTorch eager {name}: out of memory ({exc})")
            torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/vae_bench.json", "w"))
