# round 2, call F: transposed backward planes + gemm_kt + dM1 on transposed tiles + forward-only mode + spmm v2: parity first
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gemm.py -q -x -k "kt_transposed" > gpurun_out/r2f_kt.log 2>&1; rc=$?; tail -n 5 gpurun_out/r2f_kt.log
if [ $rc -ne 0 ]; then echo "GEMM_KT FAILED"; tail -n 30 gpurun_out/r2f_kt.log; fi
timeout 150 python tools/fused_check.py > gpurun_out/r2f_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2f_fused_all.log | tail -n 8; awk '/^\[/{c=$1} /grad/{ if ($4+0 > 4e-6) print c, $2, $4}' gpurun_out/r2f_fused_all.log
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2f_fused_all.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -n 12 gpurun_out/r2f_pytest_gpu.log
timeout 400 python bench.py --workload 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_b5.json 2> gpurun_out/r2f_b5.err; tail -n 3 gpurun_out/r2f_b5.err; cut -c1-300 gpurun_out/r2f_b5.json
REGT_B200_LIB=regt-gcn_b200/lib/variants/fpre/libregt_b200.so timeout 400 python bench.py --workload 5 --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_b5_fpre.json 2> gpurun_out/r2f_b5_fpre.err; cut -c1-200 gpurun_out/r2f_b5_fpre.json
timeout 200 python tools/prof_cell.py > gpurun_out/r2f_prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell_.*_f|k_wgrad_m1_kt|k_gemm_kt" -s 4 -c 4 -o gpurun_out/r2f_cell_f python tools/prof_cell.py > gpurun_out/r2f_ncu.log 2>&1; tail -n 5 gpurun_out/r2f_ncu.log
