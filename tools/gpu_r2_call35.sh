#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for i in 1 2; do
timeout 600 python bench.py --path denoise --steps 10 --no-cpu-baseline --no-eager-baseline > $O/r02ae_bench_denoise_$i.json 2> $O/r02ae_bench_denoise.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02ae_bench_denoise_$i.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
PY
done
timeout 300 python tools/unet_profile.py 14 > $O/r02ae_unet_op_profile_T14.txt 2>&1; sed -n 1,14p $O/r02ae_unet_op_profile_T14.txt
timeout 600 python tools/vae_bench.py --no-eager > $O/r02ae_vae_bench.log 2>&1; tail -4 $O/r02ae_vae_bench.log
