#!/bin/bash
# ncu evidence for one round (run under gpurun from the repo root; one GPU).  Outputs in gpurun_out/:
#   <tag>_plain_full.log        the bench line of the command that is profiled (no profiler attached)
#   <tag>_launches_full.csv     every launch of that command with its device time (cold, serialised)
#   <tag>_dram_full.csv         DRAM bytes + time of the layer kernels at full size
#   <tag>_med_set_full.ncu-rep  --set full (+ source) of one step's six layer-kernel launches, N=400k graph
tag=${1:-r2a}
FULL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch-metric --no-ssl-metric"
MED="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch-metric --no-ssl-metric --nodes 400000 --raw-edges 5000000"
$FULL > gpurun_out/${tag}_plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches_full.csv $FULL > gpurun_out/${tag}_ncu_a.log 2>&1
$FULL > /dev/null 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_disga -c 12 --csv --log-file gpurun_out/${tag}_dram_full.csv $FULL > gpurun_out/${tag}_ncu_b.log 2>&1
$MED > gpurun_out/${tag}_plain_med.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_disga -s 18 -c 6 -f -o gpurun_out/${tag}_med_set_full $MED > gpurun_out/${tag}_ncu_c.log 2>&1
ls -la gpurun_out | tail -8
