#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
EVW_VAE_DEBUG=1 timeout 300 python tools/vae_debug.py encode > $O/r02m_vae_debug_encode.log 2>&1; echo "rc=$?"; grep -n "nonfinite [1-9]\|finite:\|rror" $O/r02m_vae_debug_encode.log | head -8; head -30 $O/r02m_vae_debug_encode.log
EVW_VAE_DEBUG=1 timeout 300 python tools/vae_debug.py decode > $O/r02m_vae_debug_decode.log 2>&1; echo "rc=$?"; grep -n "nonfinite [1-9]\|finite:\|rror" $O/r02m_vae_debug_decode.log | head -8; head -30 $O/r02m_vae_debug_decode.log
