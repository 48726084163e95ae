#!/bin/bash
# CLIP ViT image encoder: parity vs transformers
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_clip.py -x -q -s > $O/r02q_clip_tests.log 2>&1; echo "clip tests rc=$?"; grep -E "rel L2|passed|failed|Error|error" $O/r02q_clip_tests.log | tail -20
