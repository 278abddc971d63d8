set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py tests/test_gpu_gemm.py -q > gpurun_out/pytest_gemm.log 2>&1; tail -8 gpurun_out/pytest_gemm.log
