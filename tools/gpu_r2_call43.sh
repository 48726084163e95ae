#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vggt.py tests/test_gpu_resize.py tests/test_gpu_clip.py tests/test_gpu_gemm.py -x -q -m gpu -s > $O/r02al_tests.log 2>&1
echo "tests rc=$?"
grep -E "vggt|passed|failed|Error|error" $O/r02al_tests.log | tail -20
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --no-eager --out $O/r02al_vggt_bench_S25.json > $O/r02al_vggt_bench_S25.log 2>&1
echo "bench25 rc=$?"; tail -1 $O/r02al_vggt_bench_S25.log
timeout 600 python tools/vggt_bench.py --frames 49 --steps 2 --no-eager --no-point-head --out $O/r02al_vggt_bench_S49.json > $O/r02al_vggt_bench_S49.log 2>&1
echo "bench49 rc=$?"; tail -1 $O/r02al_vggt_bench_S49.log
