mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -n 2 gpurun_out/bench_default.err
timeout 300 python bench.py --precision tf32x3 --workload 4 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b4_tf32.json 2> gpurun_out/b4_tf32.err; tail -n 2 gpurun_out/b4_tf32.err
timeout 300 python bench.py --precision tf32x3 --workload 5 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b5_tf32_n1.json 2> gpurun_out/b5_tf32_n1.err; tail -n 2 gpurun_out/b5_tf32_n1.err
python - <<'PY'
import json
for f in ("bench_default","b4_tf32","b5_tf32_n1"):
    try: d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    except Exception as e: print(f,"ERR",e); continue
    print(f, round(d["value"],1), round(d["ms_per_step"],3), "mb", d["config"].get("micro_batch"), "e2e", round(d["e2e"]["value"],1))
    print("   roofline", {k: (round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k!="peak_source"})
    print("   fp32 mode", d.get("fp32_parity_mode")); print("   cpu", d.get("cpu_baseline"))
PY
