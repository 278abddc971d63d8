set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q -k "peer or nccl" > gpurun_out/pytest_shard.log 2>&1; tail -n 5 gpurun_out/pytest_shard.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tools/microbench/exchange_ab.py > gpurun_out/exchange_ab_n$N.log 2>&1; grep "floats" gpurun_out/exchange_ab_n$N.log || tail -n 20 gpurun_out/exchange_ab_n$N.log
for xp in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --exchange $xp --no-cpu-baseline > gpurun_out/b2_bf16_n${N}_$xp.json 2> gpurun_out/b2_bf16_n${N}_$xp.err; tail -n 2 gpurun_out/b2_bf16_n${N}_$xp.err
done
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for xp in (f"n{N}_peer",f"n{N}_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/b2_bf16_{xp}.json").read().strip().splitlines()[-1])
        print(xp, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["parallelism"], d["config"]["exchange_in_graph"])
    except Exception as e: print(xp, "ERR", e)
PY
