mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["kernels"]
print(sys.argv[1], round(d["value"]), "samples/s", round(d["ms_per_step"]*1000,1), "us; e2e", round(d["e2e"]["value"]), {n: v["ms_per_step"] for n, v in k.items()})
PY
}
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_new.json 2>gpurun_out/f_new.err; show gpurun_out/f_new.json
REGT_FEAT_SCALAR=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_scalar.json 2>gpurun_out/f_scalar.err; show gpurun_out/f_scalar.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_new.json 2>gpurun_out/f_new.err; show gpurun_out/f_new.json
