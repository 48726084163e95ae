#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 75 python -m pytest tests/test_gpu_vggt.py -x -q -m gpu -s -k reference_modules > $O/r02ay_vggt_1b_golden.log 2>&1; echo "rc=$?"
grep -E "vggt-1b|passed|failed|Error" $O/r02ay_vggt_1b_golden.log | tail -5
