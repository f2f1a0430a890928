#!/bin/bash
# round 2, session p (1 GPU): ncu launch list of the bench command, long enough
# to contain the timed steps after the segment tuner's candidate launches
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 1 --headline-only > $O/r02p_bench_headline.json 2> $O/r02p_bench_headline.err; echo "plain exit $?"; cut -c1-200 $O/r02p_bench_headline.json
timeout 900 ncu -k regex:soda --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02p_bench_launches.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/r02p_ncu_list.log 2>&1; echo "ncu list exit $?"
grep -c soda_stream $O/r02p_bench_launches.csv
