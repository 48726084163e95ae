#!/bin/bash
# Round-1 eleventh GPU pass: 16-warp GEGLU epilogue — parity, micro-benchmark, step A/B.
O=gpurun_out/c11; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "geglu" > $O/gemm_bench_w16.log
EVW_GEMM_GEGLU_WARPS=8 timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "geglu" > $O/gemm_bench_w8.log
timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_w16.log 2>&1
EVW_GEMM_GEGLU_WARPS=8 timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_w8.log 2>&1
ls -la $O
