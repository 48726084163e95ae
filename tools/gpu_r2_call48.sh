#!/bin/bash
# ncu --set full of the VGGT global attention launch and of the VGGT elementwise kernels; summaries written on the box
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:spatial_attn8_kernel --launch-skip 25 --launch-count 2 -o $O/r02aq_ncu_vggt_attention -f python tools/vggt_bench.py --frames 25 --profile-once --no-point-head > $O/r02aq_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
python tools/ncu_summary.py $O/r02aq_ncu_vggt_attention.ncu-rep > $O/r02aq_ncu_full_vggt_attention.txt 2>&1; head -40 $O/r02aq_ncu_full_vggt_attention.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:qknorm_rope_kernel|bilinear_ac_kernel|relu_inplace_kernel" --launch-skip 47 --launch-count 10 -o $O/r02aq_ncu_vggt_elem -f python tools/vggt_bench.py --frames 25 --profile-once --no-point-head > $O/r02aq_ncu_elem.log 2>&1; echo "ncu elem rc=$?"
python tools/ncu_summary.py $O/r02aq_ncu_vggt_elem.ncu-rep > $O/r02aq_ncu_full_vggt_elem.txt 2>&1
rm -f $O/r02aq_ncu_vggt_elem.ncu-rep
ls -la $O
