# round 2, call O: SpMM v4b (two edges per trip, LDS.128, batched staging loads), CTAs-per-SM sweep
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "spmm or plan or feat or golden" > gpurun_out/r2o_pytest_spmm.log 2>&1; tail -n 2 gpurun_out/r2o_pytest_spmm.log
for c in 2 3 4 1; do
REGT_SPMM_CTAS=$c timeout 200 python tools/prof_spmm.py > gpurun_out/r2o_spmm_c$c.log 2>&1; echo "ctas=$c"; tail -n 1 gpurun_out/r2o_spmm_c$c.log
REGT_SPMM_CTAS=$c PB=64 timeout 200 python tools/prof_spmm.py > gpurun_out/r2o_spmm_b64_c$c.log 2>&1; tail -n 1 gpurun_out/r2o_spmm_b64_c$c.log
done
REGT_SPMM_CTAS=3 timeout 300 python -m pytest tests -m gpu -q -k "spmm" > gpurun_out/r2o_pytest_spmm3.log 2>&1; tail -n 2 gpurun_out/r2o_pytest_spmm3.log
REGT_SPMM_CTAS=4 timeout 300 python -m pytest tests -m gpu -q -k "spmm" > gpurun_out/r2o_pytest_spmm4.log 2>&1; tail -n 2 gpurun_out/r2o_pytest_spmm4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_spmm" -s 3 -c 1 -f -o gpurun_out/r2o_spmm python tools/prof_spmm.py > gpurun_out/r2o_spmm_ncu.log 2>&1; tail -n 3 gpurun_out/r2o_spmm_ncu.log
