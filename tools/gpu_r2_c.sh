# round 2, call C: GPU test suite on the fused 3xTF32 path, A/B of the bring-up case, cfg5 / cfg4 bench lines
set -x
mkdir -p gpurun_out
CASE=3 timeout 300 python tools/fused_check.py > gpurun_out/r2c_case3_fused.log 2>&1; tail -n 30 gpurun_out/r2c_case3_fused.log
CASE=3 REGT_UNFUSED=1 timeout 300 python tools/fused_check.py > gpurun_out/r2c_case3_unfused.log 2>&1; tail -n 30 gpurun_out/r2c_case3_unfused.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -n 25 gpurun_out/r2c_pytest_gpu.log
timeout 600 python bench.py --workload 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_b5.json 2> gpurun_out/r2c_b5.err; tail -n 3 gpurun_out/r2c_b5.err; cut -c1-300 gpurun_out/r2c_b5.json
timeout 600 python bench.py --workload 4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2c_b4.json 2> gpurun_out/r2c_b4.err; tail -n 3 gpurun_out/r2c_b4.err; cut -c1-300 gpurun_out/r2c_b4.json
