#!/bin/bash
# conf_select with 7 launches + warp-aggregated histograms; GEGLU epilogue with prefetched tensor-memory loads
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_reproj.py tests/test_gpu_gemm.py -x -q > $O/r02t_tests.log 2>&1; echo "reproj+gemm tests rc=$?"; tail -3 $O/r02t_tests.log
timeout 300 python tools/secondary_bench.py 2>&1 | grep -i "conf_select" | tee $O/r02t_conf_select.log
for e in 1 0; do echo "== EVW_GEMM_GEGLU_PREFETCH=$e"; EVW_GEMM_GEGLU_PREFETCH=$e timeout 300 python tools/gemm_bench.py 2>&1 | grep -i geglu; done | tee $O/r02t_geglu_prefetch.log
timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02t_bench_denoise.json 2> $O/r02t_bench_denoise.err; echo "bench rc=$?"
EVW_GEMM_GEGLU_PREFETCH=0 timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02t_bench_denoise_noprefetch.json 2> $O/r02t_bench_noprefetch.err; echo "bench noprefetch rc=$?"
python - <<'PY'
import json
for f in ("r02t_bench_denoise.json", "r02t_bench_denoise_noprefetch.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    except Exception as e:
        print(f, "ERR", e)
PY
