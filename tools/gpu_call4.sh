#!/bin/bash
# Round-1 fourth GPU pass: v6 attention parity + timing, denoise launch list (own kernels only), secondary kernels.
O=gpurun_out/c4; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
timeout 500 python tools/attn_bench.py > $O/attn_bench.log 2>&1
timeout 300 python tools/reproj_bench.py > $O/reproj_bench.log 2>&1
timeout 300 python tools/secondary_bench.py > $O/secondary_bench.log 2>&1
for v in 0 5 7; do
  EVW_ATTN_V5=$v timeout 300 python bench.py --path denoise --no-cpu-baseline > $O/bench_denoise_v$v.log 2>&1
done
EVW_ATTN_V5=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spatial_attn' -s 1 -c 1 -o $O/full_attn6 \
    python tools/ncu_gemm.py attn > $O/ncu_full_attn6.log 2>&1
K='tc_gemm|spatial_attn|temporal_attn|gn_stats|gn_finalize|gn_apply|layer_norm|upsample2x|downsplit|pre_kernel|post_kernel|silu_f16|cast_f16|fill_f32|timestep_embed|add_f32|nchw_to|nhwc_to'
EVW_ATTN_V5=5 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 4000 --csv --log-file $O/launches_denoise.csv \
  python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_denoise.log 2>&1
ls -la $O
