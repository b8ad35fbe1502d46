#!/bin/bash
G=/tmp/graph_A.npy
run() { # name plan proj3x
  EDIS_AT_PLAN=$2 EDIS_PROJ3X=$3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --graph-cache $G 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernels']
print('$1', 'ms/step %.1f' % d['ms_per_step'], 'Medges/s %.1f' % (d['value']/1e6), ' '.join('%s=%.1f(%.0f)' % (a.replace('disga_',''), b['ms_per_step'], b['gbs']) for a, b in k.items()))
"
}
run proj_3x proj 1
run proj_plain proj 0
