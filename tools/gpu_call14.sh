#!/bin/bash
# Round-1 extra pass: pinned-buffer e2e test, BASELINE config-5 panorama size (1024x2048 -> 128x256 latents), one step.
O=gpurun_out/c14; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_reproj.py -m gpu -x -q ) > $O/pytest_gpu_reproj.log 2>&1
timeout 600 python bench.py --path denoise --pano-height 1024 --pano-width 2048 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_denoise_1024x2048.log 2>&1
ls -la $O
