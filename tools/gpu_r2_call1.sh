#!/bin/bash
# round 2, call 1: parity at 1e-3 (small / full width / BASELINE shapes), full GPU suite, bench with the eager baseline,
# split-precision A/B, splat point-order experiment, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q -s > $O/r02a_unet_parity.log 2>&1; echo "unet parity rc=$?" | tee -a $O/r02a_rc.txt
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_unet.py > $O/r02a_gpu_suite.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/r02a_rc.txt
timeout 600 python bench.py > $O/r02a_bench_n1.json 2> $O/r02a_bench_n1.err; echo "bench rc=$?" | tee -a $O/r02a_rc.txt
EVW_UNET_SPLIT=0 timeout 300 python bench.py --path denoise --no-cpu-baseline --no-eager-baseline > $O/r02a_bench_denoise_nosplit.json 2> $O/r02a_bench_nosplit.err; echo "bench nosplit rc=$?" | tee -a $O/r02a_rc.txt
timeout 300 python tools/splat_sort_experiment.py > $O/r02a_splat_sort_experiment.log 2>&1; echo "splat sort rc=$?" | tee -a $O/r02a_rc.txt
timeout 300 python __graft_entry__.py --smoke > $O/r02a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02a_rc.txt
tail -5 $O/r02a_unet_parity.log $O/r02a_gpu_suite.log $O/r02a_splat_sort_experiment.log $O/r02a_smoke.log
head -c 1500 $O/r02a_bench_n1.json
