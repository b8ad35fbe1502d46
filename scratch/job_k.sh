mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trainers.py tests/test_gpu_bundled.py -m gpu -q -x -k "not accuracy" > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/r2_pytest7.log | head -8)
for r in 1 0; do
EDIS_RING_PAIR=$r timeout 300 python bench.py --steps 3 --warmup 3 --no-epoch-metric --no-cpu-baseline > gpurun_out/r2_bench_pair$r.log 2> gpurun_out/r2_bench_pair$r.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_pair$r.log").read().strip().splitlines()[-1]); print("ring_pair=$r", round(d["ms_per_step"],1), d["secondary"]["supedge_step"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2_bench_pair$r.err").read()[-800:])
PY
done
