#!/bin/bash
# Round-1 ninth GPU pass: final evidence with the default configuration (smoke, bench both arms, launch lists, ncu captures).
O=gpurun_out/c9; mkdir -p $O
( time timeout 600 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
( time timeout 600 python bench.py ) > $O/bench_n1.log 2>&1
( time timeout 400 python bench.py --impl reference ) > $O/bench_ref.log 2>&1
timeout 300 python tools/secondary_bench.py > $O/secondary_bench.log 2>&1
timeout 300 python tools/reproj_bench.py > $O/reproj_bench.log 2>&1
K='tc_gemm|spatial_attn|temporal_attn|gn_stats|gn_finalize|gn_apply|layer_norm|upsample2x|downsplit|pre_kernel|post_kernel|silu_f16|cast_f16|fill_f32|timestep_embed|add_f32|nchw_to|nhwc_to'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 4000 --csv --log-file $O/launches_denoise.csv \
  python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_denoise.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'select_|compact_|cube_splat|resolve_multi|pack_points|lift_' -c 2000 --csv --log-file $O/launches_reproj.csv \
  python bench.py --path reproj --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_reproj.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spatial_attn' -s 1 -c 1 -o $O/full_attn8 python tools/ncu_gemm.py attn > $O/ncu_full_attn8.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'cube_splat|resolve_multi' -s 12 -c 2 -o $O/full_reproj \
  python bench.py --path reproj --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_reproj.log 2>&1
ls -la $O
