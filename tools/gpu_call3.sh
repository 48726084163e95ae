#!/bin/bash
# Round-1 third GPU pass: parity (all attention variants, splat flag combinations), variant sweeps, secondary kernels.
O=gpurun_out/c3; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
timeout 500 python tools/attn_bench.py > $O/attn_bench.log 2>&1
timeout 300 python tools/reproj_bench.py > $O/reproj_bench.log 2>&1
timeout 300 python tools/secondary_bench.py > $O/secondary_bench.log 2>&1
for v in 5 7; do
  EVW_ATTN_V5=$v timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_v$v.log 2>&1
done
EVW_ATTN_V5=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spatial_attn' -s 1 -c 1 -o $O/full_attn6 \
    python tools/ncu_gemm.py attn > $O/ncu_full_attn6.log 2>&1
timeout 300 python tools/unet_profile.py 14 > $O/unet_profile_T14.log 2>&1
ls -la $O
