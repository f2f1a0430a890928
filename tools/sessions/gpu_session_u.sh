#!/bin/bash
set -x
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu19.log 2>&1; tail -3 $O/pytest_gpu19.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke4.log 2>&1; tail -1 $O/smoke4.log
python bench.py --steps 20 --warmup 3 > $O/bench_r1j.json 2> $O/bench_r1j.err; cat $O/bench_r1j.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-400 $O/bench_ref.json
true
