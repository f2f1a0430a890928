#!/bin/bash
# round 2, session ai (1 GPU): ncu launch list of the bench command for the
# final library (the list of session ah had the segment tuner, timing its
# candidates under the profiler, settle on 6 segments: 632 us per launch)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 1 --headline-only > $O/r02ai_bench_headline.json 2> $O/r02ai_bench_headline.err; echo "bench exit $?"; cut -c1-200 $O/r02ai_bench_headline.json
for i in 1 2; do
timeout 600 ncu -k regex:soda --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02ai_bench_launches_$i.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/r02ai_ncu_list_$i.log 2>&1; echo "ncu list exit $?"
python - <<PY
import csv
rows=list(csv.reader(open('$O/r02ai_bench_launches_$i.csv')))
hdr=None; data=[]
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
vals=[round(float(d['Metric Value'].replace(',',''))/1000) for d in data if d['Metric Name']=='gpu__time_duration.sum']
print(len(vals), vals[-22:], data[-2]['Grid Size'])
PY
done
