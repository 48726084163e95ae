#!/bin/bash
# Round-1 seventh GPU pass: GEMM cluster / weight-multicast mode — parity, micro-benchmark, step A/B.
O=gpurun_out/c7; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q ) > $O/pytest_gpu_gemm.log 2>&1
timeout 300 python tools/gemm_bench.py > $O/gemm_bench.log 2>&1
EVW_GEMM_CLUSTER=1 timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_cluster1.log 2>&1
EVW_GEMM_CLUSTER=0 timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_cluster0.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gemm.py ) > $O/pytest_gpu_rest.log 2>&1
EVW_GEMM_CLUSTER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm' -s 1 -c 1 -o $O/full_conv_cluster python tools/ncu_gemm.py conv > $O/ncu_conv1.log 2>&1
EVW_GEMM_CLUSTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm' -s 1 -c 1 -o $O/full_conv_plain python tools/ncu_gemm.py conv > $O/ncu_conv0.log 2>&1
timeout 200 python tools/secondary_bench.py > $O/secondary_bench.log 2>&1
ls -la $O
