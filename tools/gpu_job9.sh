set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_gemm.log 2>&1; tail -n 5 gpurun_out/pytest_gemm.log
timeout 240 python tools/microbench/gemm_ab.py > gpurun_out/gemm_ab.log 2>&1; tail -n 20 gpurun_out/gemm_ab.log
