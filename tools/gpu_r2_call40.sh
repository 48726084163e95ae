#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for m in 1024 960 1300 640 1024 960; do
  EVW_GEMM_PAIR_MIN_K=$m timeout 600 python bench.py --path denoise --steps 10 --no-cpu-baseline --no-eager-baseline > $O/r02ai_bench_pair_min_k_$m.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02ai_bench_pair_min_k_$m.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("EVW_GEMM_PAIR_MIN_K=$m", round(d["value"], 3), round(d["ms_per_step"], 2), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a in ("tc_gemm_kernel",)})
PY
done
