mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/r2_pytest4.log | head -20)
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-epoch-metric --no-cpu-baseline --no-ssl-metric > gpurun_out/r2_sweep_$name.log 2> gpurun_out/r2_sweep_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_sweep_$name.log").read().strip().splitlines()[-1]); print("$name", round(d["ms_per_step"],1), {k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/r2_sweep_$name.err").read()[-800:])
PY
}
run base X=1
for v in ns3 ns6 ns8fwd4 rpw1 rpw16 cost2 cost16; do
  [ -f variants/$v.so ] && run $v EDIS_LIB=variants/$v.so
done
