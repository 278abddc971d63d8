mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tools/microbench/exchange_ab.py > gpurun_out/exchange_ab_n$N.log 2>&1; grep "floats" gpurun_out/exchange_ab_n$N.log || tail -n 20 gpurun_out/exchange_ab_n$N.log
