#!/bin/bash
# DPT heads with 32 frames per pass (was 8): VGGT tests + bench at both chunk sizes
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vggt.py -x -q -m gpu > $O/r02au_vggt_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r02au_vggt_tests.log; grep -E "^FAILED|Error|assert " $O/r02au_vggt_tests.log | head
for c in 8 32; do
  timeout 600 python tools/vggt_bench.py --frames 25 --steps 3 --no-eager --dpt-chunk $c --out $O/r02au_vggt_bench_S25_chunk$c.json > $O/r02au_vggt_bench_S25_chunk$c.log 2>&1; echo "chunk $c rc=$?"; tail -1 $O/r02au_vggt_bench_S25_chunk$c.log | cut -c150-600
done
