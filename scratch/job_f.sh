# usage: job_f.sh <ngpus>
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}.log 2> gpurun_out/r2_bench_n${N}.err; echo "bench A N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_n${N}.err)
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n${N}.log").read().strip().splitlines()[-1])
    print("A N=$N", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["setup_s"]); print(d["breakdown"]); print({k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()}); print(d["secondary"]); print(d["config"]["partition"])
except Exception as e: print("ERR", e)
PY
(timeout 600 $TR --master-port 29531 tests/dist_check_gpu.py --out gpurun_out/r2_dist_check_n${N}.json > gpurun_out/r2_dist_n${N}.log 2>&1; echo "dist rc=$?"; tail -2 gpurun_out/r2_dist_n${N}.log | cut -c1-1500)
if [ "$2" = "B" ]; then
for gnn in AT GCN SAGE; do
  (timeout 420 $TR --master-port 29551 bench.py --gpus $N --config B --gnn_type $gnn --steps 3 --warmup 3 --no-ssl-metric > gpurun_out/r2_bench_B_${gnn}_n${N}.log 2> gpurun_out/r2_bench_B_${gnn}_n${N}.err; echo "bench B $gnn N=$N rc=$?"; tail -c 600 gpurun_out/r2_bench_B_${gnn}_n${N}.err)
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_B_${gnn}_n${N}.log").read().strip().splitlines()[-1])
    print("B $gnn N=$N", d["value"], d["ms_per_step"], d["config"]["edges"], d["config"]["setup_s"], d["config"]["partition"]); print(d["breakdown"])
except Exception as e: print("ERR", e)
PY
done
fi
nvidia-smi --query-gpu=memory.used --format=csv | head -3
