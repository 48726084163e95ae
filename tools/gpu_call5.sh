#!/bin/bash
# Round-1 fifth GPU pass: v7 attention (P in tensor memory) parity + timing, e2e with pinned buffers, launch list.
O=gpurun_out/c5; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_unet_ops.py -m gpu -x -q ) > $O/pytest_gpu_ops.log 2>&1
timeout 500 python tools/attn_bench.py > $O/attn_bench.log 2>&1
for v in 9 10 11; do
  EVW_ATTN_V5=$v timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_v$v.log 2>&1
done
EVW_ATTN_V5=9 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spatial_attn' -s 1 -c 1 -o $O/full_attn7 \
    python tools/ncu_gemm.py attn > $O/ncu_full_attn7.log 2>&1
timeout 300 python bench.py --path reproj > $O/bench_reproj.log 2>&1
timeout 300 python tools/secondary_bench.py > $O/secondary_bench.log 2>&1
ls -la $O
