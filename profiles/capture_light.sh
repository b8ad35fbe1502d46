#!/bin/bash
# Reduced ncu evidence (GPU budget): DRAM bytes of one step's six layer-kernel launches at full size + --set full of the
# same six on the med graph.  The launch list of a whole step is capture.sh's (tag r2a).
tag=${1:-r2b}
FULL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch-metric --no-ssl-metric"
MED="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch-metric --no-ssl-metric --nodes 400000 --raw-edges 5000000"
$FULL > gpurun_out/${tag}_plain_full.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_disga -s 18 -c 6 --csv --log-file gpurun_out/${tag}_dram_full.csv $FULL > gpurun_out/${tag}_ncu_b.log 2>&1
$MED > gpurun_out/${tag}_plain_med.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_disga -s 18 -c 6 -f -o gpurun_out/${tag}_med_set_full $MED > gpurun_out/${tag}_ncu_c.log 2>&1
ls -la gpurun_out | grep ${tag}
