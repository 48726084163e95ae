#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for c in qkv to_out geglu; do
  EVW_GEMM_CLUSTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o $O/r02i_full_$c python tools/ncu_gemm.py $c > $O/r02i_ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spatial_attn -s 2 -c 1 -o $O/r02i_full_attn python tools/ncu_gemm.py attn > $O/r02i_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
