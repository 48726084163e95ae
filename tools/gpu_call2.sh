#!/bin/bash
# Round-1 second GPU pass: parity of the new kernels, variant micro-benchmarks, bench, ncu evidence.
O=gpurun_out/c2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
timeout 400 python tools/attn_bench.py > $O/attn_bench.log 2>&1
timeout 300 python tools/reproj_bench.py > $O/reproj_bench.log 2>&1
( time timeout 600 python bench.py ) > $O/bench_n1.log 2>&1
timeout 300 python tools/unet_profile.py 14 > $O/unet_profile_T14.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_denoise.csv \
  python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_denoise.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_reproj.csv \
  python bench.py --path reproj --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_reproj.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'cube_splat|resolve_multi' -s 12 -c 2 -o $O/full_reproj \
  python bench.py --path reproj --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_reproj.log 2>&1
for c in attn gn; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spatial_attn|gn_stats|gn_apply' -s 1 -c 2 -o $O/full_$c \
    python tools/ncu_gemm.py $c > $O/ncu_full_$c.log 2>&1
done
ls -la $O
