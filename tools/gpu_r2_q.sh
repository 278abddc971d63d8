# round 2, call Q: SpMM v4c (pointer loop, next block prefetched into L2)
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_plan_spmm.py -m gpu -q -x > gpurun_out/r2q_pytest_spmm.log 2>&1; tail -n 3 gpurun_out/r2q_pytest_spmm.log
timeout 200 python tools/prof_spmm.py > gpurun_out/r2q_spmm.log 2>&1; tail -n 1 gpurun_out/r2q_spmm.log
PB=64 timeout 200 python tools/prof_spmm.py > gpurun_out/r2q_spmm_b64.log 2>&1; tail -n 1 gpurun_out/r2q_spmm_b64.log
PW=4 PB=64 timeout 200 python tools/prof_spmm.py > gpurun_out/r2q_spmm_w4.log 2>&1; tail -n 1 gpurun_out/r2q_spmm_w4.log
PPART=0 PB=64 timeout 200 python tools/prof_spmm.py > gpurun_out/r2q_spmm_b64_nopart.log 2>&1; tail -n 1 gpurun_out/r2q_spmm_b64_nopart.log
REGT_SPMM_CTAS=2 PB=64 timeout 200 python tools/prof_spmm.py > gpurun_out/r2q_spmm_b64_c2.log 2>&1; tail -n 1 gpurun_out/r2q_spmm_b64_c2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_spmm" -s 3 -c 1 -f -o gpurun_out/r2q_spmm python tools/prof_spmm.py > gpurun_out/r2q_spmm_ncu.log 2>&1; tail -n 3 gpurun_out/r2q_spmm_ncu.log
