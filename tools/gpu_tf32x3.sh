mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
for wl in 2 3 4; do
timeout 300 python bench.py --precision tf32x3 --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b${wl}_tf32.json 2> gpurun_out/b${wl}_tf32.err; tail -n 2 gpurun_out/b${wl}_tf32.err
done
python - <<'PY'
import json
for wl in (2,3,4):
    try: d=json.loads(open(f"gpurun_out/b{wl}_tf32.json").read().strip().splitlines()[-1])
    except Exception as e: print(wl,"ERR",e); continue
    print(wl, round(d["value"],1), round(d["ms_per_step"],3), "mb", d["config"].get("micro_batch"))
    for k,v in list(d["kernels"].items())[:12]: print("   ",k,v)
PY
