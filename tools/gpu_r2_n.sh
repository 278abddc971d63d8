# round 2, call N: SpMM v4 (generic-pointer edge entries): parity tests, timing, ncu
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "spmm or plan or feat or golden" > gpurun_out/r2n_pytest_spmm.log 2>&1; tail -n 4 gpurun_out/r2n_pytest_spmm.log
timeout 200 python tools/prof_spmm.py > gpurun_out/r2n_spmm_plain.log 2>&1; tail -n 1 gpurun_out/r2n_spmm_plain.log
PB=64 timeout 200 python tools/prof_spmm.py > gpurun_out/r2n_spmm_b64.log 2>&1; tail -n 1 gpurun_out/r2n_spmm_b64.log
PW=4 PB=64 timeout 200 python tools/prof_spmm.py > gpurun_out/r2n_spmm_w4.log 2>&1; tail -n 1 gpurun_out/r2n_spmm_w4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_spmm" -s 3 -c 1 -f -o gpurun_out/r2n_spmm python tools/prof_spmm.py > gpurun_out/r2n_spmm_ncu.log 2>&1; tail -n 3 gpurun_out/r2n_spmm_ncu.log
