#!/bin/bash
# tuning sweep on the GPU box: same graph (cached), each variant 3 steps
G=/tmp/graph_A.npy
run() { # name lib kv
  EDIS_LIB=$2 EDIS_KV=$3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --graph-cache $G 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernel_ms']
print('$1', 'ms/step %.1f' % d['ms_per_step'], 'Medges/s %.1f' % (d['value']/1e6), ' '.join('%s=%.1f' % (a.replace('disga_',''), b) for a, b in k.items()))
"
}
run kv4u1b2 "" 0
run kv2u2b2 "" 2
run kv4u2b1 variants/libedis_kv4u2b1.so 0
run kv4u2b2 variants/libedis_kv4u2b2.so 0
