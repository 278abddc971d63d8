# multi-GPU evidence run: bash tools/gpu_scale.sh N   (under gpurun --gpus N)
set -x
N=${1:-8}
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 300 29621 bench.py --gpus $N --no-cpu-baseline > gpurun_out/b2_bf16_n${N}_peer.json 2> gpurun_out/b2_bf16_n${N}_peer.err; tail -n 2 gpurun_out/b2_bf16_n${N}_peer.err
run 300 29622 bench.py --gpus $N --no-cpu-baseline --exchange nccl > gpurun_out/b2_bf16_n${N}_nccl.json 2> gpurun_out/b2_bf16_n${N}_nccl.err; tail -n 2 gpurun_out/b2_bf16_n${N}_nccl.err
run 600 29623 bench.py --gpus $N --workload 5 --precision tf32x3 --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/b5_tf32_n${N}.json 2> gpurun_out/b5_tf32_n${N}.err; tail -n 2 gpurun_out/b5_tf32_n${N}.err
run 300 29624 tools/microbench/exchange_ab.py > gpurun_out/exchange_ab_n$N.log 2>&1; grep "floats" gpurun_out/exchange_ab_n$N.log || tail -n 20 gpurun_out/exchange_ab_n$N.log
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in (f"b2_bf16_n{N}_peer", f"b2_bf16_n{N}_nccl", f"b5_tf32_n{N}"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["scaling"], round(d["value"],1), d["unit"], round(d["ms_per_step"],4), "ms; e2e", round(d["e2e"]["value"],1), "|", d["config"]["parallelism"])
    except Exception as e: print(f, "ERR", e)
PY
