#!/bin/bash
# round 2, session n (8 GPUs): NUMA placement diagnostics
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02n_topo.txt 2>&1; head -14 $O/r02n_topo.txt | cut -c1-200
for g in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "gpu $g $b $(cat /sys/bus/pci/devices/$b/numa_node 2>&1)"; done
timeout 600 python tools/numa_diag.py > $O/r02n_numa_diag.jsonl 2> $O/r02n_numa_diag.err; cat $O/r02n_numa_diag.jsonl | cut -c1-900; tail -3 $O/r02n_numa_diag.err
