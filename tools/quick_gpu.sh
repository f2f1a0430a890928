timeout 600 python tools/bench_configs.py run jacobi2d seidel2d 2>&1 | cut -c1-300 | tail -4
timeout 600 python tools/tune2d.py run 2>&1 | python -c "
import sys, json
rows=[json.loads(l) for l in sys.stdin if l.startswith('{')]
best={}
for r in rows:
    if r['tb'] not in best or r['gcell_per_s']>best[r['tb']]['gcell_per_s']: best[r['tb']]=r
for tb in sorted(best): print(best[tb])
"
