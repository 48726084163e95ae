#!/bin/bash
# VAE: full test file (incl. 576x1024 vs oracle), pipeline with the native VAE, full-size benchmark
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_pipeline.py -x -q -s > $O/r02n_vae_tests.log 2>&1; echo "vae+pipeline tests rc=$?"; grep -E "rel L2|passed|failed|Error|error|decode plan" $O/r02n_vae_tests.log | tail -30
timeout 900 python tools/vae_bench.py > $O/r02n_vae_bench.log 2>&1; echo "vae bench rc=$?"; tail -12 $O/r02n_vae_bench.log
