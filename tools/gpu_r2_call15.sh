#!/bin/bash
# config-5 size on one GPU: 3-clip iterative episode at 1024x2048, 25 frames, 50 steps (memory + time check before the 8-GPU run)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python bench.py --path iterative --pano-height 1024 --pano-width 2048 --iter-steps 50 --iter-warmup-steps 2 > $O/r02p_bench_iterative_config5_n1.json 2> $O/r02p_bench_iterative_config5_n1.err; echo "rc=$?"
tail -5 $O/r02p_bench_iterative_config5_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02p_bench_iterative_config5_n1.json").read().strip().splitlines()[-1])
    print(d["value"], d.get("ms_per_episode"), d.get("ms_per_stage_per_episode"), d.get("finite_output"), d.get("memory_points_per_segment"))
    print(d.get("clocks"), d["config"])
except Exception as e:
    print("ERR", e)
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
