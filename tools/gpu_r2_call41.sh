#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_vggt.py -x -q -m gpu -s > $O/r02aj_vggt_tests.log 2>&1
echo "tests rc=$?"
grep -E "vggt|passed|failed|Error|error" $O/r02aj_vggt_tests.log | tail -30
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --out $O/r02aj_vggt_bench_S25.json > $O/r02aj_vggt_bench_S25.log 2>&1
echo "bench rc=$?"; tail -3 $O/r02aj_vggt_bench_S25.log
