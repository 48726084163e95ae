#!/bin/bash
# Round-1 eighth GPU pass: elect.sync MMA/TMA issue (no uniform-operand waterfall) — parity, GEMM micro-benchmark, step A/B.
O=gpurun_out/c8; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
timeout 300 python tools/gemm_bench.py > $O/gemm_bench.log 2>&1
timeout 300 python tools/attn_bench.py --new-only > $O/attn_bench.log 2>&1
EVW_GEMM_CLUSTER=0 timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_cluster0.log 2>&1
EVW_GEMM_CLUSTER=1 timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_cluster1.log 2>&1
EVW_GEMM_CLUSTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm' -s 1 -c 1 -o $O/full_conv_plain python tools/ncu_gemm.py conv > $O/ncu_conv0.log 2>&1
timeout 300 python -m pytest tests/test_gpu_unet.py -m gpu -q -s 2>&1 | grep -E "rel L2|passed|failed" > $O/unet_parity.log
ls -la $O
