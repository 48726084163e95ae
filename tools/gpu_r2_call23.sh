#!/bin/bash
# A/B: the GEMMs with and without their output stores (EVW_DEBUG_NO_STORE build) on the level-0 shapes
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
echo "== normal"; timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "plain|cluster" | grep -E "L0|L1" | tee $O/r02x_gemm_bench_normal.log
echo "== no stores"; EVW_LIB=$PWD/gpurun_variants/libevw_nostore.so timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "plain|cluster" | grep -E "L0|L1" | tee $O/r02x_gemm_bench_nostore.log
