# validate: full GPU suite, smoke, tf32x3 benches of cfg 2/3/4
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 5 gpurun_out/smoke.log
for wl in 2 3; do
timeout 200 python bench.py --precision tf32x3 --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b${wl}_tf32.json 2> gpurun_out/b${wl}_tf32.err; tail -n 2 gpurun_out/b${wl}_tf32.err
done
timeout 300 python bench.py --precision tf32x3 --workload 4 --micro-batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b4_tf32.json 2> gpurun_out/b4_tf32.err; tail -n 2 gpurun_out/b4_tf32.err
python - <<'PY'
import json
for wl in (2,3,4):
    try: d=json.loads(open(f"gpurun_out/b{wl}_tf32.json").read().strip().splitlines()[-1])
    except Exception as e: print(wl,"ERR",e); continue
    print(wl, d["value"], d["ms_per_step"], d["roofline_step"])
    for k,v in d["kernels"].items(): print("   ",k,v)
PY
