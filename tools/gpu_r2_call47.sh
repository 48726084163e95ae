#!/bin/bash
# VGGT: new tests (loader, reference call sequence through dropin), ncu --set full of the global attention and the VGGT elementwise kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vggt.py tests/test_gpu_resize.py -x -q -m gpu > $O/r02ap_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02ap_tests.log
grep -E "Error|assert" $O/r02ap_tests.log | head -10
# launches of spatial_attn8_kernel in one forward: 24 DINOv2 + (frame, global) x 24 -> skip 25 = the first global attention, then the next frame attention
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spatial_attn8_kernel --launch-skip 25 --launch-count 2 -o $O/r02ap_ncu_vggt_attention -f python tools/vggt_bench.py --frames 25 --profile-once --no-point-head > $O/r02ap_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:qknorm_rope_kernel|bilinear_ac_kernel|layer_norm_kernel|relu_inplace_kernel" --launch-skip 50 --launch-count 40 -o $O/r02ap_ncu_vggt_elem -f python tools/vggt_bench.py --frames 25 --profile-once --no-point-head > $O/r02ap_ncu_elem.log 2>&1; echo "ncu elem rc=$?"
ls -la $O/*.ncu-rep
