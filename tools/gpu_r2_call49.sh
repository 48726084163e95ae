#!/bin/bash
# verification after the q/k-norm + bilinear rewrites: full GPU suite, smoke, VGGT bench + launch list, default bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02ar_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02ar_rc.txt; tail -3 $O/r02ar_tests.log
python -c "from __graft_entry__ import smoke; smoke()" > $O/r02ar_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02ar_rc.txt; tail -5 $O/r02ar_smoke.log
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --no-eager --out $O/r02ar_vggt_bench_S25.json > $O/r02ar_vggt_bench_S25.log 2>&1; echo "vggt bench rc=$?"; tail -1 $O/r02ar_vggt_bench_S25.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02ar_vggt_launches.csv python tools/vggt_bench.py --frames 25 --profile-once --no-point-head > $O/r02ar_vggt_ncu.log 2>&1
python tools/launch_summary.py $O/r02ar_vggt_launches.csv --title "VGGT-1B forward, 25 frames 392x518, depth head only (final kernels)" > $O/r02ar_vggt_launch_summary.txt 2>&1; head -12 $O/r02ar_vggt_launch_summary.txt; gzip -f $O/r02ar_vggt_launches.csv
T1=$(date +%s)
timeout 1500 python bench.py --no-cpu-baseline > $O/r02ar_bench_n1.json 2> $O/r02ar_bench_n1.err; echo "bench rc=$? in $(( $(date +%s) - T1 )) s" | tee -a $O/r02ar_rc.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02ar_bench_n1.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"))
    it = d.get("iterative"); print("  iterative", it and (it["value"], it["ms_per_episode"], it["finite_output"], it["ms_per_stage_per_episode"].get("vggt"), it.get("vggt_forward")))
    print("  clocks", d.get("clocks"))
except Exception as e:
    print("ERR", e)
PY
