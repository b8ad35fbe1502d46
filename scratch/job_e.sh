mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/r2_pytest5.log | head -20)
(timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_n1.err)
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print({k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()}); print(d["roofline"]["frac"], d["cpu_baseline"]); print(d["secondary"])
PY
