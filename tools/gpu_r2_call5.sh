#!/bin/bash
# round 2, call 5: relaxed-wait placement, PointMemory tests, iterative episode bench (both modes), full default bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02e_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02e_rc.txt
tail -3 $O/r02e_tests.log
timeout 300 python tools/secondary_bench.py > $O/r02e_secondary_bench.log 2>&1; echo "secondary bench rc=$?" | tee -a $O/r02e_rc.txt
timeout 600 python bench.py --path iterative --iter-mode incremental > $O/r02e_bench_iterative.json 2> $O/r02e_bench_iterative.err; echo "iter rc=$?" | tee -a $O/r02e_rc.txt
timeout 600 python bench.py --path iterative --iter-mode reference > $O/r02e_bench_iterative_reference_mode.json 2> $O/r02e_bench_iterative_ref.err; echo "iter ref-mode rc=$?" | tee -a $O/r02e_rc.txt
timeout 900 python bench.py > $O/r02e_bench_n1.json 2> $O/r02e_bench_n1.err; echo "bench rc=$?" | tee -a $O/r02e_rc.txt
grep -i "equi\|conf\|lift" $O/r02e_secondary_bench.log
tail -3 $O/r02e_bench_iterative.err
python - <<'PY'
import json
for f in ("r02e_bench_iterative.json", "r02e_bench_iterative_reference_mode.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["config"])
        print("   ", {k: d.get(k) for k in ("ms_per_episode", "ms_per_stage_per_episode", "memory_points_per_segment")})
    except Exception as e:
        print(f, "ERR", e)
try:
    d = json.loads(open("gpurun_out/r02e_bench_n1.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    print("  eager", {a: (round(b["ms_per_step"], 1) if isinstance(b, dict) else b) for a, b in d["gpu_eager_baseline"].items() if a != "how"})
    r = d["reproj"]; print("  reproj", r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"]["frac"])
    it = d.get("iterative"); print("  iterative", it and (it["value"], it["ms_per_episode"], it["ms_per_stage_per_episode"]))
except Exception as e:
    print("bench ERR", e)
PY
