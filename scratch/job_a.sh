mkdir -p gpurun_out
(timeout 300 ./scratch/gemm_emul_bin > gpurun_out/r2_gemm_emul.log 2>&1; cat gpurun_out/r2_gemm_emul.log)
(timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/r2_pytest3.log | head -20)
for cfg in "1 1 1" "0 0 0" "1 0 0" "0 1 0" "0 0 1"; do
  set -- $cfg
  EDIS_RING_FWD=$1 EDIS_RING_DST=$2 EDIS_RING_SRC=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-epoch-metric --no-cpu-baseline --no-ssl-metric > gpurun_out/r2_bench_ring_$1$2$3.log 2> gpurun_out/r2_bench_ring_$1$2$3.err
  echo "ring fwd=$1 dst=$2 src=$3 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_ring_$1$2$3.log").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"]); print({k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2_bench_ring_$1$2$3.err").read()[-1500:])
PY
done
