#!/bin/bash
# final closing run of the round: full GPU suite, smoke, VGGT-1B at 25 (with the eager legs) and 49 frames, default bench (all paths)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02av_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02av_rc.txt; tail -3 $O/r02av_tests.log
python -c "from __graft_entry__ import smoke; smoke()" > $O/r02av_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02av_rc.txt; tail -5 $O/r02av_smoke.log
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --out $O/r02av_vggt_bench_S25.json > $O/r02av_vggt_bench_S25.log 2>&1; echo "vggt S25 rc=$?"; tail -1 $O/r02av_vggt_bench_S25.log | cut -c150-900
timeout 600 python tools/vggt_bench.py --frames 49 --steps 2 --no-point-head --out $O/r02av_vggt_bench_S49.json > $O/r02av_vggt_bench_S49.log 2>&1; echo "vggt S49 rc=$?"; tail -1 $O/r02av_vggt_bench_S49.log | cut -c150-900
T1=$(date +%s)
timeout 1500 python bench.py > $O/r02av_bench_n1.json 2> $O/r02av_bench_n1.err; echo "bench rc=$? in $(( $(date +%s) - T1 )) s" | tee -a $O/r02av_rc.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02av_bench_n1.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"))
    r = d["reproj"]; print("  reproj", r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"]["frac"])
    it = d.get("iterative"); print("  iterative", it and (it["value"], it["ms_per_episode"], it["finite_output"], it["ms_per_stage_per_episode"].get("vggt"), it.get("vggt_forward", {}).get("ms")))
    print("  clocks", d.get("clocks"))
except Exception as e:
    print("ERR", e)
PY
