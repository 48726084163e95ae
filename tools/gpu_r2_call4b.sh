#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for v in A B C; do
  EVW_LIB=$PWD/evoworld_b200/_lib/variants/lib_$v.so timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -2 > $O/r02d_variant_$v.log; echo "variant $v: $(tail -1 $O/r02d_variant_$v.log)"
done
timeout 300 python -m pytest tests/test_gpu_gemm.py -q 2>&1 | tail -15
