set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_shard.py -x -q > gpurun_out/pytest_tc.log 2>&1; tail -15 gpurun_out/pytest_tc.log
timeout 300 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/b2_bf16_v5.json 2> gpurun_out/b2_bf16_v5.err; tail -3 gpurun_out/b2_bf16_v5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b2_bf16_v5.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k,v in d["kernels"].items(): print("  ",k,v)
PY
