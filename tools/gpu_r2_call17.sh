#!/bin/bash
# BASELINE config 5 on 8 GPUs: 3-clip iterative episode x 8 scenes (one per rank), 1024x2048, 25 frames, 50 steps,
# NCCL all-gather of the latents at every clip boundary
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --path iterative --pano-height 1024 --pano-width 2048 --iter-steps 50 --iter-warmup-steps 2 > $O/r02r_bench_iterative_config5_n8.json 2> $O/r02r_bench_iterative_config5_n8.err; echo "rc=$?"
tail -3 $O/r02r_bench_iterative_config5_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02r_bench_iterative_config5_n8.json").read().strip().splitlines()[-1])
    print(d["value"], d["n_gpus"], d.get("ms_per_episode"), d.get("ms_per_stage_per_episode"), d.get("finite_output"))
    print(d.get("clocks"))
except Exception as e:
    print("ERR", e)
PY
