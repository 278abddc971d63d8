# multi-GPU runs (launch exactly as the driver does); $1 = number of GPUs
set -x
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload 5 --precision tf32x3 --micro-batch 8 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b5_tf32_n1.json 2> gpurun_out/b5_tf32_n1.err; tail -2 gpurun_out/b5_tf32_n1.err
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N > gpurun_out/b2_bf16_n$N.json 2> gpurun_out/b2_bf16_n$N.err; tail -2 gpurun_out/b2_bf16_n$N.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --workload 5 --precision tf32x3 --steps 3 --warmup 3 --no-graph > gpurun_out/b5_tf32_n$N.json 2> gpurun_out/b5_tf32_n$N.err; tail -2 gpurun_out/b5_tf32_n$N.err
  true
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b*_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["scaling"], round(d["value"],1), "samples/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(f, "ERR", e)
PY
