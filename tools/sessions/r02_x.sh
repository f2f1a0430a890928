#!/bin/bash
# round 2, session x (1 GPU): what bounds the e2e pipeline now that compute is
# hidden at every chunk count - the cost of cutting a transfer into pieces
# (tools/probe/copy_gap_probe.py), and the ramped layouts with the new tuner
# threshold
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 400 python tools/probe/copy_gap_probe.py > $O/r02x_copy_gap_probe.jsonl 2> $O/r02x_copy_gap_probe.err
python - <<PY
import json
for l in open('$O/r02x_copy_gap_probe.jsonl'):
  d = json.loads(l)
  print(d['directions'], '%4d MiB' % d['piece_mib'], 'events' if d['events'] else '      ', '2 streams' if d['two_streams'] else '         ', '%.2f ms  %.1f GB/s' % (d['ms_best'], d['gbs']))
PY
tail -3 $O/r02x_copy_gap_probe.err
timeout 400 python tools/e2e_ab.py 16 -16 -20 -24 20 24 -32 16 > $O/r02x_e2e_ramped.jsonl 2> $O/r02x_e2e_ramped.err
cat $O/r02x_e2e_ramped.jsonl; tail -2 $O/r02x_e2e_ramped.err
