#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for e in "X=0" "EVW_GEMM_GEGLU_WARPS=16" "EVW_GEMM_PLAIN_WARPS=16"; do echo "== $e"; env $e timeout 300 python tools/gemm_bench.py 2>&1 | grep plain | grep -E "qkv|geglu|linear"; done
env EVW_GEMM_GEGLU_WARPS=16 EVW_GEMM_PLAIN_WARPS=16 timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -2
env EVW_GEMM_GEGLU_WARPS=16 EVW_GEMM_PLAIN_WARPS=16 timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02j_bench_denoise_16w.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02j_bench_denoise_16w.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
PY
