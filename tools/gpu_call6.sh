#!/bin/bash
# Round-1 sixth GPU pass: v8 attention parity + timing, full GPU test suite with the new default (v7), full bench line.
O=gpurun_out/c6; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
timeout 400 python tools/attn_bench.py --new-only > $O/attn_bench.log 2>&1
EVW_ATTN_V5=14 timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_v14.log 2>&1
( time timeout 600 python bench.py ) > $O/bench_n1.log 2>&1
timeout 300 python bench.py --impl reference > $O/bench_ref.log 2>&1
timeout 300 python bench.py --path denoise --frames 25 --steps 3 --no-cpu-baseline > $O/bench_denoise_T25.log 2>&1
timeout 300 python tools/secondary_bench.py > $O/secondary_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spatial_attn' -s 1 -c 1 -o $O/full_attn7 \
    python tools/ncu_gemm.py attn > $O/ncu_full_attn7.log 2>&1
ls -la $O
