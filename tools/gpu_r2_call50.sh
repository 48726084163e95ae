#!/bin/bash
# patchify kernel: VGGT tests + bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vggt.py -x -q -m gpu > $O/r02as_vggt_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02as_vggt_tests.log; grep -E "Error|assert " $O/r02as_vggt_tests.log | head
timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --no-eager --out $O/r02as_vggt_bench_S25.json > $O/r02as_vggt_bench_S25.log 2>&1; echo "vggt bench rc=$?"; tail -1 $O/r02as_vggt_bench_S25.log
