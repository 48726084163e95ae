#!/bin/bash
# round 2, call 4: lean epilogue + fastdiv + suspend-hint waits + whole-plan CUDA graph
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q > $O/r02d_gemm_tests.log 2>&1; rc=$?; echo "gemm tests rc=$rc" | tee $O/r02d_rc.txt
tail -3 $O/r02d_gemm_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python tools/gemm_bench.py > $O/r02d_gemm_bench.log 2>&1; echo "gemm bench rc=$?" | tee -a $O/r02d_rc.txt
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_unet_ops.py -x -q -s > $O/r02d_unet_tests.log 2>&1; echo "unet tests rc=$?" | tee -a $O/r02d_rc.txt
timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02d_bench_denoise.json 2> $O/r02d_bench_denoise.err; echo "bench rc=$?" | tee -a $O/r02d_rc.txt
EVW_UNET_GRAPH=0 timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02d_bench_denoise_nograph.json 2> $O/r02d_bench_nograph.err; echo "bench nograph rc=$?" | tee -a $O/r02d_rc.txt
timeout 300 python tools/unet_profile.py 14 > $O/r02d_unet_op_profile_T14.txt 2>&1; echo "profile rc=$?" | tee -a $O/r02d_rc.txt
cat $O/r02d_gemm_bench.log
grep -E "rel L2|passed|failed|Error" $O/r02d_unet_tests.log | tail -14
head -14 $O/r02d_unet_op_profile_T14.txt
python - <<'PY'
import json
for f in ("r02d_bench_denoise.json", "r02d_bench_denoise_nograph.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["config"].get("graph_replays"), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    except Exception as e:
        print(f, "ERR", e)
PY
