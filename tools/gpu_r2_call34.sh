#!/bin/bash
# fused nearest-x2 + 3x3 convolution (four 2x2 phase GEMMs with strided TMA stores): parity and effect
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q -k "upsample or tma_store or pair_and" > $O/r02ad_gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -12 $O/r02ad_gemm_tests.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_vae.py -x -q -s > $O/r02ad_unet_tests.log 2>&1; echo "unet/vae tests rc=$?"; grep -E "rel L2|passed|failed|Error" $O/r02ad_unet_tests.log | tail -22 | cut -c1-200
timeout 600 python bench.py --path denoise --steps 10 --no-cpu-baseline --no-eager-baseline > $O/r02ad_bench_denoise.json 2> $O/r02ad_bench_denoise.err; echo "bench rc=$?"
timeout 600 python tools/vae_bench.py --no-eager > $O/r02ad_vae_bench.log 2>&1; tail -4 $O/r02ad_vae_bench.log
timeout 300 python tools/unet_profile.py 14 > $O/r02ad_unet_op_profile_T14.txt 2>&1; sed -n 1,14p $O/r02ad_unet_op_profile_T14.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02ad_bench_denoise.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
PY
