# round 2, call B: bring-up of the fused 3xTF32 cell (short timeouts: a deadlocked mbarrier must not hold the box),
# then the GPU test suite and the cfg5 / cfg4 bench lines
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
FWD_ONLY=1 CASE=0 timeout 120 python tools/fused_check.py > gpurun_out/r2b_fused_fwd0.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/r2b_fused_fwd0.log
FWD_ONLY=1 timeout 180 python tools/fused_check.py > gpurun_out/r2b_fused_fwd.log 2>&1; echo "rc=$?"; tail -n 8 gpurun_out/r2b_fused_fwd.log
timeout 300 python tools/fused_check.py > gpurun_out/r2b_fused_all.log 2>&1; echo "rc=$?"; tail -n 80 gpurun_out/r2b_fused_all.log
REGT_UNFUSED=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu_unfused.log 2>&1; tail -n 15 gpurun_out/r2b_pytest_gpu_unfused.log
REGT_UNFUSED=1 timeout 600 python bench.py --workload 5 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_b5_unfused.json 2> gpurun_out/r2b_b5_unfused.err; tail -n 3 gpurun_out/r2b_b5_unfused.err; cut -c1-600 gpurun_out/r2b_b5_unfused.json
