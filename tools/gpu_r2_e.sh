# round 2, call E: 16 epilogue warps + balanced setmaxnreg split + M1 cache: parity first (short timeout: a deadlock must
# not hold the box), then the suite, the cfg5 bench, A/B variants, ncu of the fused kernels, the MN-major tf32 layout probe
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2e_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v grad gpurun_out/r2e_fused_all.log | tail -n 8
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2e_fused_all.log; exit 1; fi
timeout 60 python tools/umma_probe.py > gpurun_out/r2e_umma_probe.log 2>&1; tail -n 14 gpurun_out/r2e_umma_probe.log
for v in tn2 tn4; do
  CASE=3 REGT_B200_LIB=regt-gcn_b200/lib/variants/$v/libregt_b200.so timeout 150 python tools/fused_check.py > gpurun_out/r2e_case3_$v.log 2>&1
  grep -E "linear1|_attention|linear_z.weight|^\[" gpurun_out/r2e_case3_$v.log
done
CASE=3 timeout 150 python tools/fused_check.py > gpurun_out/r2e_case3_base.log 2>&1; grep -E "linear1|_attention|linear_z.weight|^\[" gpurun_out/r2e_case3_base.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -n 6 gpurun_out/r2e_pytest_gpu.log
timeout 400 python bench.py --workload 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_b5.json 2> gpurun_out/r2e_b5.err; tail -n 3 gpurun_out/r2e_b5.err; cut -c1-300 gpurun_out/r2e_b5.json
for v in fpre tn2; do
  REGT_B200_LIB=regt-gcn_b200/lib/variants/$v/libregt_b200.so timeout 400 python bench.py --workload 5 --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2e_b5_$v.json 2> gpurun_out/r2e_b5_$v.err; cut -c1-200 gpurun_out/r2e_b5_$v.json
done
timeout 200 python tools/prof_cell.py > gpurun_out/r2e_prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell_.*_f|k_wgrad_m1_pm|k_gemm_tn_tma" -s 5 -c 5 -o gpurun_out/r2e_cell_f python tools/prof_cell.py > gpurun_out/r2e_ncu.log 2>&1; tail -n 5 gpurun_out/r2e_ncu.log
