#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu13.log 2>&1; tail -5 $O/pytest_gpu13.log
python tools/bench_configs.py run blur sobel2d erosion xcorr contrast denoise2d denoise3d > $O/configs_v7.log 2> $O/configs_v7.err; cat $O/configs_v7.log
