#!/bin/bash
# closing ncu evidence of round 2: launch list of the bench command (denoise + reprojection), --set full of the GEMM with the
# GroupNorm-statistics epilogue (3x3 conv + row vector at level 0), of the GroupNorm apply kernel, of a VAE decoder conv
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
EVW_UNET_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm|spatial_attn|gn_|layer_norm|temporal_attn|upsample|downsplit|pre_kernel|post_kernel|silu|timestep|cast_f16|fill_f32|set_step' -c 3000 --csv --log-file $O/r02s_denoise_launches.csv python bench.py --path denoise --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/r02s_ncu_denoise.log 2>&1; echo "ncu denoise list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'splat|resolve|select_|compact_|lift|pack' -c 600 --csv --log-file $O/r02s_reproj_launches.csv python bench.py --path reproj --steps 2 --warmup 3 --no-cpu-baseline > $O/r02s_ncu_reproj.log 2>&1; echo "ncu reproj list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o $O/r02s_full_conv_gnstats python tools/ncu_gemm.py conv_stats > $O/r02s_ncu_conv_gnstats.log 2>&1; echo "ncu conv+stats rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o $O/r02s_full_conv python tools/ncu_gemm.py conv_rv > $O/r02s_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 3 -c 1 -o $O/r02s_full_gn_apply python tools/ncu_gemm.py gn > $O/r02s_ncu_gn.log 2>&1; echo "ncu gn rc=$?"
gzip -f $O/r02s_denoise_launches.csv $O/r02s_reproj_launches.csv
ls -la $O | grep r02s
