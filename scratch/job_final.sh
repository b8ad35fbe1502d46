mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/r2_pytest_final.log | head -10)
(timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.log 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?")
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1_final.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["setup_s"]); print({k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()}); print(d["roofline"]["frac"], d["roofline"]["traffic"], d["gpu_launches"], d["clocks"]); print(d["cpu_baseline"]["value"], d["secondary"]["supedge_step"]["ms_per_step"], d["secondary"]["cora_full_epoch_ms"]["reference_rng_sampler_device_metrics"])
PY
