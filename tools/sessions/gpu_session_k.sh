#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu12.log 2>&1; tail -5 $O/pytest_gpu12.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke2.log 2>&1; tail -2 $O/smoke2.log
python tools/bench_configs.py run contrast erosion xcorr > $O/configs_v6.log 2> $O/configs_v6.err; cat $O/configs_v6.log
