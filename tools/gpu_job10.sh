mkdir -p gpurun_out
echo ---- main
timeout 200 python tools/microbench/tn_debug.py > gpurun_out/tn_debug.log 2>&1; grep -v "best hyp\|B-side" gpurun_out/tn_debug.log | cut -c1-150
timeout 200 python tools/microbench/tn_debug.py > gpurun_out/tn_debug.log 2>&1; grep -v "best hyp\|B-side" gpurun_out/tn_debug.log | cut -c1-150
timeout 180 python -m pytest tests/test_gpu_gemm.py -q > gpurun_out/pytest_gemm.log 2>&1; tail -n 5 gpurun_out/pytest_gemm.log
timeout 240 python tools/microbench/gemm_ab.py > gpurun_out/gemm_ab.log 2>&1; grep "tn x3" gpurun_out/gemm_ab.log
echo ---- late
REGT_B200_LIB=$PWD/regt-gcn_b200/lib/variants/tn_late/libregt_b200.so timeout 240 python tools/microbench/gemm_ab.py > gpurun_out/gemm_ab_late.log 2>&1; grep "tn x3" gpurun_out/gemm_ab_late.log
