#!/bin/bash
# closing run after the fused up-sampling convolution: full GPU tests, smoke, default bench, reference arm, T = 25 / 1024x2048
# lines, denoise launch list
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02af_tests.log 2>&1; echo "tests rc=$?" | tee $O/r02af_rc.txt
tail -3 $O/r02af_tests.log
timeout 600 python __graft_entry__.py --smoke > $O/r02af_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02af_rc.txt; tail -4 $O/r02af_smoke.log
timeout 1200 python bench.py > $O/r02af_bench_n1.json 2> $O/r02af_bench_n1.err; echo "bench rc=$?" | tee -a $O/r02af_rc.txt
timeout 600 python bench.py --path denoise --frames 25 --no-cpu-baseline --no-eager-baseline > $O/r02af_bench_denoise_T25.json 2> $O/r02af_bench_T25.err; echo "bench T25 rc=$?" | tee -a $O/r02af_rc.txt
timeout 600 python bench.py --path denoise --pano-height 1024 --pano-width 2048 --no-cpu-baseline --no-eager-baseline > $O/r02af_bench_denoise_1024x2048.json 2> $O/r02af_bench_1024.err; echo "bench 1024x2048 rc=$?" | tee -a $O/r02af_rc.txt
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02af_bench_reference.json 2> $O/r02af_bench_reference.err; echo "reference arm rc=$?" | tee -a $O/r02af_rc.txt
EVW_UNET_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm|spatial_attn|gn_|layer_norm|temporal_attn|upsample|downsplit|pre_kernel|post_kernel|silu|timestep|cast_f16|fill_f32|set_step' -c 3000 --csv --log-file $O/r02af_denoise_launches.csv python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/r02af_ncu_denoise.log 2>&1; echo "ncu denoise list rc=$?"
gzip -f $O/r02af_denoise_launches.csv
python - <<'PY'
import json
def show(f):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = (d.get("roofline") or {}).get("kernels") or {}
        print(f, d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("gpu_launches"),
              {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
        return d
    except Exception as e:
        print(f, "ERR", e)
d = show("r02af_bench_n1.json")
if d:
    print("  eager", {a: (round(b["ms_per_step"], 1) if isinstance(b, dict) else b) for a, b in d["gpu_eager_baseline"].items() if a != "how"})
    r = d["reproj"]; print("  reproj", r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"]["frac"], r["roofline"].get("frac_with_a11"))
    it = d.get("iterative"); print("  iterative", it and (it["value"], it["ms_per_episode"], it["ms_per_stage_per_episode"], it["finite_output"]))
    print("  clocks", d.get("clocks"))
show("r02af_bench_denoise_T25.json"); show("r02af_bench_denoise_1024x2048.json"); show("r02af_bench_reference.json")
PY
