#!/bin/bash
# adaptive tile shape (choose_bx): full GPU suite, VGGT bench, denoise bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02at_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02at_rc.txt; tail -3 $O/r02at_tests.log; grep -E "^FAILED|Error|assert " $O/r02at_tests.log | head
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --no-eager --out $O/r02at_vggt_bench_S25.json > $O/r02at_vggt_bench_S25.log 2>&1; echo "vggt bench rc=$?"; tail -1 $O/r02at_vggt_bench_S25.log
timeout 600 python bench.py --path denoise --steps 10 --no-cpu-baseline --no-eager-baseline > $O/r02at_bench_denoise.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02at_bench_denoise.json").read().strip().splitlines()[-1])
print("denoise", round(d["value"], 3), round(d["ms_per_step"], 2), d["roofline"]["frac"])
PY
