#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_resize.py tests/test_gpu_gemm.py -x -q -s > $O/r02ac_tests.log 2>&1; echo "resize+gemm tests rc=$?"; grep -E "panoramas|passed|failed|Error" $O/r02ac_tests.log | tail -5
timeout 600 python bench.py --path iterative > $O/r02ac_bench_iterative.json 2> $O/r02ac_bench_iterative.err; echo "iter rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02ac_bench_iterative.json").read().strip().splitlines()[-1])
print(d["value"], d.get("ms_per_episode"), d.get("ms_per_stage_per_episode"), d.get("finite_output"), d.get("gpu_launches"))
PY
