#!/bin/bash
# tuning sweep helper: runs bench.py under env-var variants, one summary line each
G=/tmp/graph_A.npy
run() { # name, then VAR=val ...
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-epoch-metric --graph-cache $G 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernels']
print('$name', 'ms/step %.1f' % d['ms_per_step'], 'Medges/s %.1f' % (d['value']/1e6), ' '.join('%s=%.1f(%.0f/%.0f)' % (a.replace('disga_',''), b['ms_per_step'], b['gbs'], b['moved_gbs']) for a, b in k.items()))
"
}
run base X=1
run src1 EDIS_HOT_MIN_SRC=1
run src0 EDIS_HOT_MIN_SRC=0
run dst1 EDIS_HOT_MIN_DST=1
run src3 EDIS_HOT_MIN_SRC=3
