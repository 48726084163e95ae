#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python tools/unet_profile.py 14 > $O/r02w_unet_op_profile_T14.txt 2>&1; echo "rc=$?"
sed -n 1,14p $O/r02w_unet_op_profile_T14.txt; grep -A45 "GEMM shapes" $O/r02w_unet_op_profile_T14.txt
