mkdir -p gpurun_out
T="tests/test_gpu_trainers.py -m gpu -q -k train_steps"
for cfg in "default:X=1" "ring0:EDIS_RING=0" "exact:EDIS_LIB=variants/exact.so" "exact_ring0:EDIS_LIB=variants/exact.so EDIS_RING=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== $name"; env $envs timeout 300 python -m pytest $T 2>&1 | grep -E "passed|failed|AssertionError"
done
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-epoch-metric --no-cpu-baseline --no-ssl-metric > gpurun_out/r2_sweep2_$name.log 2> gpurun_out/r2_sweep2_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_sweep2_$name.log").read().strip().splitlines()[-1]); print("$name", round(d["ms_per_step"],1), {k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/r2_sweep2_$name.err").read()[-800:])
PY
}
run base X=1
for v in rpw1 rpw2 ns3 exact; do run $v EDIS_LIB=variants/$v.so; done
