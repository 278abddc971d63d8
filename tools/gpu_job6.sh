set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py tests/test_gpu_gemm.py -q > gpurun_out/pytest_gemm.log 2>&1; tail -25 gpurun_out/pytest_gemm.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_umma.py > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
