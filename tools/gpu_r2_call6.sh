#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py -x -q > $O/r02h_gemm_tests.log 2>&1; echo "gemm tests rc=$?" | tee $O/r02h_rc.txt
timeout 300 python tools/gemm_bench.py > $O/r02h_gemm_bench.log 2>&1; echo "gemm bench rc=$?" | tee -a $O/r02h_rc.txt
timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02h_bench_denoise.json 2> $O/r02h_bench_denoise.err; echo "bench rc=$?" | tee -a $O/r02h_rc.txt
tail -2 $O/r02h_gemm_tests.log; grep -E "plain|cluster" $O/r02h_gemm_bench.log | grep -E "res|linear|qkv"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02h_bench_denoise.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
PY
