# round 2, call AS: bench.py with the spmm key timed at the kernel (regt_profile events) as well as through the wrapper
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2as_b5.json 2> gpurun_out/r2as_b5.err; tail -n 2 gpurun_out/r2as_b5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2as_b5.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['spmm']))
PY
