#!/bin/bash
# round 2, call 2: cta_group::2 GEMM — parity first (bounded), then micro-bench, then the step bench + pipeline tests + smoke
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -s > $O/r02b_gemm_tests.log 2>&1; rc=$?; echo "gemm tests rc=$rc" | tee $O/r02b_rc.txt
tail -5 $O/r02b_gemm_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python tools/gemm_bench.py > $O/r02b_gemm_bench.log 2>&1; echo "gemm bench rc=$?" | tee -a $O/r02b_rc.txt
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_unet_ops.py -x -q -s > $O/r02b_unet_pipeline_tests.log 2>&1; echo "unet+pipeline tests rc=$?" | tee -a $O/r02b_rc.txt
timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02b_bench_denoise.json 2> $O/r02b_bench_denoise.err; echo "bench rc=$?" | tee -a $O/r02b_rc.txt
EVW_GEMM_CLUSTER=0 timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02b_bench_denoise_1cta.json 2> $O/r02b_bench_denoise_1cta.err; echo "bench 1cta rc=$?" | tee -a $O/r02b_rc.txt
timeout 300 python __graft_entry__.py --smoke > $O/r02b_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02b_rc.txt
cat $O/r02b_gemm_bench.log
tail -12 $O/r02b_unet_pipeline_tests.log; tail -3 $O/r02b_smoke.log
python - <<'PY'
import json
for f in ("r02b_bench_denoise.json", "r02b_bench_denoise_1cta.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    except Exception as e:
        print(f, "ERR", e)
PY
