#!/bin/bash
# round 2, call 3: auto pair-mode threshold sweep on the step, ncu of the epilogue-bound K=320 GEMMs, pipeline tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_pipeline.py -x -q > $O/r02c_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02c_rc.txt
for k in 512 1024 2048; do
  EVW_GEMM_PAIR_MIN_K=$k timeout 300 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02c_bench_denoise_mink$k.json 2> $O/r02c_bench_mink$k.err; echo "bench mink=$k rc=$?" | tee -a $O/r02c_rc.txt
done
for c in qkv to_out geglu; do
  EVW_GEMM_CLUSTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o $O/r02c_full_$c python tools/ncu_gemm.py $c > $O/r02c_ncu_$c.log 2>&1; echo "ncu $c rc=$?" | tee -a $O/r02c_rc.txt
done
tail -4 $O/r02c_tests.log
python - <<'PY'
import json
for k in (512, 1024, 2048):
    f = f"r02c_bench_denoise_mink{k}.json"
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        kk = d["roofline"]["kernels"]
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in kk.items() if a != "how"})
    except Exception as e:
        print(f, "ERR", e)
PY
