mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bundled.py -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/r2_pytest6.log | head -8)
timeout 300 python bench.py --steps 8 --warmup 3 --no-epoch-metric --no-cpu-baseline --no-ssl-metric > gpurun_out/r2_bench_j.log 2> gpurun_out/r2_bench_j.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_j.log").read().strip().splitlines()[-1]); print(round(d["ms_per_step"],1), {k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2_bench_j.err").read()[-800:])
PY
