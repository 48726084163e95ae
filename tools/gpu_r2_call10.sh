#!/bin/bash
# GroupNorm finalize folded into apply; MUFU.EX2 f32 / f16 throughput probe
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_unet_ops.py tests/test_gpu_pipeline.py -x -q -s > $O/r02l_unet_tests.log 2>&1; echo "unet tests rc=$?"; grep -E "rel L2|passed|failed|GroupNorms|Error" $O/r02l_unet_tests.log | tail -20
timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02l_bench_denoise.json 2> $O/r02l_bench_denoise.err; echo "bench rc=$?"
timeout 60 tools/probes/mufu_probe > $O/r02l_mufu_probe.log 2>&1; cat $O/r02l_mufu_probe.log
python - <<'PY'
import json
for f in ("r02l_bench_denoise.json",):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("gpu_launches"), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    except Exception as e:
        print(f, "ERR", e)
PY
