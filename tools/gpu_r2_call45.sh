#!/bin/bash
# attention: barrier of the two warps sharing the rows (variant 5) vs the 8-warp group barrier (variant 0)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_unet_ops.py -x -q -m gpu > $O/r02an_attn_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r02an_attn_tests.log
timeout 600 python tools/attn_bench.py --pair > $O/r02an_attn_bench.log 2>&1; grep "^\[" $O/r02an_attn_bench.log | cut -c1-400
for v in 0 5 0 5; do
  EVW_ATTN_VARIANT=$v timeout 600 python bench.py --path denoise --steps 10 --no-cpu-baseline --no-eager-baseline > $O/r02an_bench_attn_variant_$v.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02an_bench_attn_variant_$v.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("EVW_ATTN_VARIANT=$v", round(d["value"], 3), round(d["ms_per_step"], 2), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if "attn" in a})
PY
done
