#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vggt.py tests/test_gpu_resize.py tests/test_gpu_clip.py -x -q -m gpu -s > $O/r02ak_vggt_tests.log 2>&1
echo "tests rc=$?"
grep -E "vggt|passed|failed|Error|error" $O/r02ak_vggt_tests.log | tail -20
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --out $O/r02ak_vggt_bench_S25.json > $O/r02ak_vggt_bench_S25.log 2>&1
echo "bench25 rc=$?"; tail -1 $O/r02ak_vggt_bench_S25.log
timeout 600 python tools/vggt_bench.py --frames 49 --steps 2 --no-point-head --out $O/r02ak_vggt_bench_S49.json > $O/r02ak_vggt_bench_S49.log 2>&1
echo "bench49 rc=$?"; tail -1 $O/r02ak_vggt_bench_S49.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02ak_vggt_launches.csv python tools/vggt_bench.py --frames 25 --profile-once --no-point-head > $O/r02ak_vggt_ncu.log 2>&1
echo "ncu rc=$?"; python tools/launch_summary.py $O/r02ak_vggt_launches.csv > $O/r02ak_vggt_launch_summary.txt 2>&1; head -30 $O/r02ak_vggt_launch_summary.txt
gzip -f $O/r02ak_vggt_launches.csv
timeout 900 python bench.py --path iterative --iter-warmup-steps 2 > $O/r02ak_bench_iterative.json 2> $O/r02ak_bench_iterative.err
echo "iter rc=$?"; tail -c 1500 $O/r02ak_bench_iterative.json
python -c "from __graft_entry__ import smoke; smoke()" > $O/r02ak_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $O/r02ak_smoke.log
