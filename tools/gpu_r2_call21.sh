#!/bin/bash
# vectorised conf_select passes; new tests (pair / single bit identity, VAE shape errors)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_reproj.py tests/test_gpu_gemm.py tests/test_gpu_vae.py -x -q > $O/r02v_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02v_tests.log
timeout 300 python tools/secondary_bench.py 2>&1 | grep -i "conf_select" | tee $O/r02v_conf_select.log
timeout 600 python bench.py --path reproj --no-cpu-baseline > $O/r02v_bench_reproj.json 2> $O/r02v_bench_reproj.err; echo "bench reproj rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02v_bench_reproj.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("frac_with_a11"), d["roofline"].get("ms_with_a11"))
PY
