set -x
mkdir -p gpurun_out
timeout 60 python gpurun_dbg.py bf16 fwd 2>&1 | tail -12 || { echo "FWD DBG HUNG/FAILED"; exit 1; }
timeout 60 python gpurun_dbg.py bf16 bwd 2>&1 | tail -12 || { echo "BWD DBG HUNG/FAILED"; exit 1; }
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_shard.py tests/test_gpu_model_parity.py -x -q > gpurun_out/pytest_tc.log 2>&1; tail -4 gpurun_out/pytest_tc.log
timeout 150 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/b2_bf16_v6.json 2> gpurun_out/b2_bf16_v6.err; tail -3 gpurun_out/b2_bf16_v6.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b2_bf16_v6.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k,v in d["kernels"].items(): print("  ",k,v)
PY
