#!/bin/bash
# Round-1 closing pass: smoke, full GPU suite, bench both arms, denoise launch list — everything at the final defaults.
O=gpurun_out/c13; mkdir -p $O
( time timeout 600 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
( time timeout 600 python bench.py ) > $O/bench_n1.log 2>&1
( time timeout 400 python bench.py --impl reference ) > $O/bench_ref.log 2>&1
K='tc_gemm|spatial_attn|temporal_attn|gn_stats|gn_finalize|gn_apply|layer_norm|upsample2x|downsplit|pre_kernel|post_kernel|silu_f16|cast_f16|fill_f32|timestep_embed|add_f32|nchw_to|nhwc_to'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 4000 --csv --log-file $O/launches_denoise.csv \
  python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_denoise.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'temporal_attn' -s 3 -c 1 -o $O/full_temporal \
  python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_temporal.log 2>&1
ls -la $O
