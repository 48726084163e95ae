#!/bin/bash
# VAE encode / temporal decode: first parity run
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vae.py -x -q -s > $O/r02m_vae_tests.log 2>&1; echo "vae tests rc=$?"; grep -E "rel L2|passed|failed|Error|error" $O/r02m_vae_tests.log | tail -30
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_gemm.py -x -q > $O/r02m_unet_tests.log 2>&1; echo "unet+gemm tests rc=$?"; tail -3 $O/r02m_unet_tests.log
