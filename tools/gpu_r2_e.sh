# round 2, call E: 16 epilogue warps + register split + M1 cache: parity, bench, ncu of the two fused kernels
set -x
mkdir -p gpurun_out
timeout 300 python tools/fused_check.py > gpurun_out/r2e_fused_all.log 2>&1; echo "rc=$?"; grep -v grad gpurun_out/r2e_fused_all.log | tail -n 8
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -n 6 gpurun_out/r2e_pytest_gpu.log
timeout 600 python bench.py --workload 5 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2e_b5.json 2> gpurun_out/r2e_b5.err; tail -n 3 gpurun_out/r2e_b5.err; cut -c1-300 gpurun_out/r2e_b5.json
timeout 300 python tools/prof_cell.py > gpurun_out/r2e_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cell_.*_f -s 2 -c 2 -o gpurun_out/r2e_cell_f python tools/prof_cell.py > gpurun_out/r2e_ncu.log 2>&1; tail -n 5 gpurun_out/r2e_ncu.log
