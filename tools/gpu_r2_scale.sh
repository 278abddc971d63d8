# round 2 multi-GPU evidence: bash tools/gpu_r2_scale.sh N [tests]   (under gpurun --gpus N)
# the driver's own command line (torchrun ... bench.py --gpus N): config 5, region shards, tf32x3, exchange inside the CUDA graph
set -x
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
if [ "$2" = "tests" ]; then
  timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -q -x > gpurun_out/r2_shard_tests_n${N}.log 2>&1; tail -n 6 gpurun_out/r2_shard_tests_n${N}.log
fi
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 600 29631 bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/r2_b5_n${N}.json 2> gpurun_out/r2_b5_n${N}.err; echo "rc=$?"; tail -n 3 gpurun_out/r2_b5_n${N}.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_b5_n{N}.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["scaling"], round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "|", d["run"]["parallelism"], "| graph", d["run"]["cuda_graph"], d["run"]["exchange_in_graph"])
    print("cross_rank_check:", d["cross_rank_check"])
    for k,v in list(d["kernels"].items())[:8]: print(f"  {k:24s} {v['ms_per_step']:9.3f} ms  x{v['launches']}")
except Exception as e: print("ERR", e)
PY
