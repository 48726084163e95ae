#!/bin/bash
# GroupNorm statistics from GEMM epilogues: parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q > $O/r02k_gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -3 $O/r02k_gemm_tests.log
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_unet_ops.py -x -q -s > $O/r02k_unet_tests.log 2>&1; echo "unet tests rc=$?"; grep -E "rel L2|passed|failed|GroupNorms|Error" $O/r02k_unet_tests.log | tail -20
timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02k_bench_denoise.json 2> $O/r02k_bench_denoise.err; echo "bench rc=$?"
EVW_GEMM_GN_STATS=0 timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02k_bench_denoise_nofuse.json 2> $O/r02k_bench_nofuse.err; echo "bench nofuse rc=$?"
python - <<'PY'
import json
for f in ("r02k_bench_denoise.json", "r02k_bench_denoise_nofuse.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("gpu_launches"), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    except Exception as e:
        print(f, "ERR", e)
PY
