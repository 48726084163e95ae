#!/bin/bash
# Round-1 tenth GPU pass: tensor-core temporal attention — parity, timing, step.
O=gpurun_out/c10; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_unet_ops.py tests/test_gpu_unet.py tests/test_gpu_reproj.py -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
timeout 300 python tools/attn_bench.py --new-only 2>&1 | grep -E "temporal|v8" > $O/attn_bench.log
EVW_TEMPORAL_ATTN_V1=1 timeout 300 python tools/attn_bench.py --new-only 2>&1 | grep -E "temporal" > $O/attn_bench_v1.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_n1.log 2>&1
timeout 300 python bench.py --path denoise --frames 25 --steps 3 --no-cpu-baseline > $O/bench_T25.log 2>&1
ls -la $O
