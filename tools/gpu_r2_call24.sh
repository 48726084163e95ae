#!/bin/bash
# TMA-store epilogue (UTMASTG): parity and A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q > $O/r02y_gemm_tests.log 2>&1; rc=$?; echo "gemm tests rc=$rc"; tail -15 $O/r02y_gemm_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then
  EVW_GEMM_STORE_TMA=0 timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -2
  exit 0
fi
for e in 1 0; do echo "== EVW_GEMM_STORE_TMA=$e"; EVW_GEMM_STORE_TMA=$e timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "plain|cluster"; done > $O/r02y_gemm_bench_store_tma.log 2>&1
python - <<'PY'
import re
txt = open("gpurun_out/r02y_gemm_bench_store_tma.log").read()
a, b = txt.split("== EVW_GEMM_STORE_TMA=0")
def parse(t):
    d = {}
    for l in t.splitlines():
        m = re.match(r"(L\d .*?)\s+(?:bn=\s*\d+\s+)?(cluster|plain)\s+([\d.]+) ms", l)
        if m: d[(m.group(1).strip(), m.group(2))] = float(m.group(3))
    return d
A, B = parse(a), parse(b)
for k in A:
    if k in B: print(f"{k[0]:34s} {k[1]:8s} direct {B[k]:.3f} -> TMA store {A[k]:.3f} ms ({100 * (1 - A[k] / B[k]):+.0f} %)")
PY
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_vae.py tests/test_gpu_clip.py -x -q > $O/r02y_unet_tests.log 2>&1; echo "unet/vae/clip tests rc=$?"; tail -3 $O/r02y_unet_tests.log | cut -c1-300
timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02y_bench_denoise.json 2> $O/r02y_bench_denoise.err; echo "bench rc=$?"
EVW_GEMM_STORE_TMA=0 timeout 600 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02y_bench_denoise_direct.json 2> $O/r02y_bench_direct.err; echo "bench direct rc=$?"
python - <<'PY'
import json
for f in ("r02y_bench_denoise.json", "r02y_bench_denoise_direct.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], {a: (round(b["ms"], 2) if isinstance(b, dict) else b) for a, b in k.items() if a != "how"})
    except Exception as e:
        print(f, "ERR", e)
PY
