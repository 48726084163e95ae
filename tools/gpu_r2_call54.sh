#!/bin/bash
# compute-sanitizer memcheck over the new VGGT elementwise kernels and the Pillow-exact resize (unit tests only)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_vggt.py tests/test_gpu_resize.py -x -q -m gpu -k "qknorm or bilinear or patchify or elementwise or bicubic or vggt_preprocess" > $O/r02aw_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -6 $O/r02aw_sanitizer.log | cut -c1-200
