mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loop.py -x -q > gpurun_out/pytest_loop.log 2>&1; tail -n 30 gpurun_out/pytest_loop.log
