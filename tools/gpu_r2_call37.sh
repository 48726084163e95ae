#!/bin/bash
# feed-forward over row panels (intermediate kept in L2): parity and A/B of the panel size
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q -s > $O/r02ag_unet_tests.log 2>&1; echo "unet tests rc=$?"; grep -E "rel L2|passed|failed|Error" $O/r02ag_unet_tests.log | tail -12 | cut -c1-200
for m in 0 80 40 160 0 80; do
  EVW_FF_PANEL_MB=$m timeout 600 python bench.py --path denoise --steps 10 --no-cpu-baseline --no-eager-baseline > $O/r02ag_bench_ff_panel_$m.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02ag_bench_ff_panel_$m.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("EVW_FF_PANEL_MB=$m", round(d["value"], 3), round(d["ms_per_step"], 2), d.get("gpu_launches"), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a in ("tc_gemm_kernel",)})
PY
done
