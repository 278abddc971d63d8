set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shard.py -x -q -k nccl > gpurun_out/pytest_nccl.log 2>&1; tail -5 gpurun_out/pytest_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload 3 --precision fp32 --steps 5 --warmup 3 > gpurun_out/b3_fp32_n2.json 2> gpurun_out/b3_fp32_n2.err; tail -3 gpurun_out/b3_fp32_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --precision bf16 > gpurun_out/b2_bf16_n2.json 2> gpurun_out/b2_bf16_n2.err; tail -3 gpurun_out/b2_bf16_n2.err
timeout 120 python gpurun_dbg.py bf16 fwd > gpurun_out/dbg_fwd.log 2>&1
timeout 120 python gpurun_dbg.py bf16 bwd > gpurun_out/dbg_bwd.log 2>&1
cat gpurun_out/dbg_fwd.log gpurun_out/dbg_bwd.log
