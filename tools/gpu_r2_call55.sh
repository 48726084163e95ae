#!/bin/bash
# compute-sanitizer memcheck over the small-configuration VGGT network tests (every kernel of the forward) and the GELU-epilogue GEMM
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_vggt.py tests/test_gpu_gemm.py -x -q -m gpu -k "small_config or other_shapes or run_vggt_inference or gelu_epilogue" > $O/r02aw_sanitizer_network.log 2>&1; echo "sanitizer rc=$?"
tail -6 $O/r02aw_sanitizer_network.log | cut -c1-200
