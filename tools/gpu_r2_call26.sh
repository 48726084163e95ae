#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for m in 1 2 3 0 1 0; do
  EVW_GEMM_STORE_TMA=$m timeout 600 python bench.py --path denoise --steps 10 --no-cpu-baseline --no-eager-baseline > $O/r02z_bench_store_tma_$m.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02z_bench_store_tma_$m.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("EVW_GEMM_STORE_TMA=$m", round(d["value"], 3), round(d["ms_per_step"], 2), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a in ("tc_gemm_kernel",)})
PY
done
