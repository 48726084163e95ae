#!/bin/bash
# Round-1 measurement pass (one gpurun call): GPU tests, bench both arms, ncu launch lists, ncu --set full captures.
O=gpurun_out/c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
( time timeout 600 python bench.py ) > $O/bench_n1.log 2>&1
timeout 300 python bench.py --impl reference > $O/bench_ref.log 2>&1
for G in 1 2 8; do timeout 200 python bench.py --path reproj --views-per-pass $G --no-cpu-baseline > $O/reproj_G$G.log 2>&1; done
EVW_SPLAT_PRETEST=1 timeout 200 python bench.py --path reproj --views-per-pass 4 --no-cpu-baseline > $O/reproj_G4_pretest.log 2>&1
timeout 300 python tools/attn_bench.py > $O/attn_bench.log 2>&1
timeout 300 python tools/gemm_bench.py > $O/gemm_bench.log 2>&1
timeout 300 python tools/unet_profile.py 14 > $O/unet_profile_T14.log 2>&1
# launch lists (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_denoise.csv \
  python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_denoise.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_reproj.csv \
  python bench.py --path reproj --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_reproj.log 2>&1
# --set full captures of the dominant kernels
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cube_splat -s 3 -c 1 -o $O/full_cube_splat \
  python bench.py --path reproj --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_splat.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resolve_multi -s 3 -c 1 -o $O/full_resolve \
  python bench.py --path reproj --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_resolve.log 2>&1
for c in geglu ff2 attn gn; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|spatial_attn|gn_stats|gn_apply' -s 1 -c 2 -o $O/full_$c \
    python tools/ncu_gemm.py $c > $O/ncu_full_$c.log 2>&1
done
ls -la $O
