#!/bin/bash
# the driver's multi-GPU launch of the default bench (all paths, VGGT stage in the iterative episode) at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02ax_bench_n2.json 2> $O/r02ax_bench_n2.err; echo "bench n2 rc=$? in $(( $(date +%s) - T0 )) s"
tail -2 $O/r02ax_bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02ax_bench_n2.json").read().strip().splitlines()[-1])
    print(d["value"], d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"))
    r = d["reproj"]; print("  reproj", r["value"], r["ms_per_step"], r["e2e"]["value"])
    it = d.get("iterative"); print("  iterative", it and (it["value"], it["ms_per_episode"], it["finite_output"], it["ms_per_stage_per_episode"].get("vggt")))
    print("  clocks", d.get("clocks"))
except Exception as e:
    print("ERR", e)
PY
