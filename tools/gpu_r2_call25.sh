#!/bin/bash
# timing experiment: the K = 320 / 640 GEMMs with the weight tile NOT re-loaded per tile (EVW_DEBUG_SKIP_W build, wrong results)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
echo "== normal"; EVW_GEMM_STORE_TMA=0 timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "plain" | tee $O/r02z_gemm_bench_normal.log
echo "== weights not re-loaded"; EVW_GEMM_STORE_TMA=0 EVW_LIB=$PWD/gpurun_variants/libevw_skipw.so timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "plain" | tee $O/r02z_gemm_bench_skipw.log
