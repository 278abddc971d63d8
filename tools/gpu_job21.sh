mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_ts.log 2>&1; tail -n 6 gpurun_out/pytest_ts.log
timeout 240 python tools/microbench/gemm_ab.py > gpurun_out/gemm_ab.log 2>&1; grep " tn " gpurun_out/gemm_ab.log
