#!/bin/bash
# closing run: full GPU suite, smoke, VGGT bench with the rewritten q/k-norm kernel, default bench (all paths) + CPU reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02am_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02am_rc.txt; tail -3 $O/r02am_tests.log
python -c "from __graft_entry__ import smoke; smoke()" > $O/r02am_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02am_rc.txt; tail -5 $O/r02am_smoke.log
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --out $O/r02am_vggt_bench_S25.json > $O/r02am_vggt_bench_S25.log 2>&1; echo "vggt bench rc=$?"; tail -1 $O/r02am_vggt_bench_S25.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02am_vggt_launches.csv python tools/vggt_bench.py --frames 25 --profile-once --no-point-head > $O/r02am_vggt_ncu.log 2>&1
python tools/launch_summary.py $O/r02am_vggt_launches.csv --title "VGGT-1B forward, 25 frames 392x518, depth head only (GELU epilogue, 16-byte q/k-norm + RoPE)" > $O/r02am_vggt_launch_summary.txt 2>&1; head -12 $O/r02am_vggt_launch_summary.txt; gzip -f $O/r02am_vggt_launches.csv
T1=$(date +%s)
timeout 1500 python bench.py > $O/r02am_bench_n1.json 2> $O/r02am_bench_n1.err; echo "bench rc=$? in $(( $(date +%s) - T1 )) s" | tee -a $O/r02am_rc.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02am_bench_n1.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"))
    r = d["reproj"]; print("  reproj", r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"]["frac"])
    it = d.get("iterative"); print("  iterative", it and (it["value"], it["ms_per_episode"], it["finite_output"], it["ms_per_stage_per_episode"]))
    print("  clocks", d.get("clocks"))
except Exception as e:
    print("ERR", e)
PY
T2=$(date +%s)
timeout 1500 python bench.py --impl reference --steps 2 --warmup 0 > $O/r02am_bench_reference.json 2> $O/r02am_bench_reference.err; echo "reference arm rc=$? in $(( $(date +%s) - T2 )) s" | tee -a $O/r02am_rc.txt
tail -c 600 $O/r02am_bench_reference.json
echo "total $(( $(date +%s) - T0 )) s"
