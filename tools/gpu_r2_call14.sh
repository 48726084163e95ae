#!/bin/bash
# round 2, call 14: full GPU test suite, smoke, default bench (all paths, iterative episode with the native VAE), reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02o_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02o_rc.txt
tail -3 $O/r02o_tests.log
timeout 600 python __graft_entry__.py --smoke > $O/r02o_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02o_rc.txt; tail -4 $O/r02o_smoke.log
timeout 1200 python bench.py > $O/r02o_bench_n1.json 2> $O/r02o_bench_n1.err; echo "bench rc=$?" | tee -a $O/r02o_rc.txt
tail -2 $O/r02o_bench_n1.err
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02o_bench_reference.json 2> $O/r02o_bench_reference.err; echo "reference arm rc=$?" | tee -a $O/r02o_rc.txt
tail -c 600 $O/r02o_bench_reference.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02o_bench_n1.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("gpu_launches"), {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    print("  eager", {a: (round(b["ms_per_step"], 1) if isinstance(b, dict) else b) for a, b in d["gpu_eager_baseline"].items() if a != "how"})
    r = d["reproj"]; print("  reproj", r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"])
    it = d.get("iterative"); print("  iterative", it and (it["value"], it["ms_per_episode"], it["ms_per_stage_per_episode"], it["finite_output"]))
    print("  clocks", d.get("clocks"))
except Exception as e:
    print("bench ERR", e)
PY
