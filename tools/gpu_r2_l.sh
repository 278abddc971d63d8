# round 2, call L: E0 stores interleaved with per-4-column h, head loss partials bounded; spmm ncu
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2l_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2l_fused_all.log | tail -n 8; awk '/^\[/{c=$1} /grad/{ if ($4+0 > 4e-6) print c, $2, $4}' gpurun_out/r2l_fused_all.log
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2l_fused_all.log; exit 1; fi
timeout 200 python tools/f_phases.py > gpurun_out/r2l_phases.log 2>&1; grep -A16 "backward" gpurun_out/r2l_phases.log
timeout 400 python bench.py --workload 5 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2l_b5.json 2> gpurun_out/r2l_b5.err; tail -n 3 gpurun_out/r2l_b5.err; cut -c1-300 gpurun_out/r2l_b5.json
timeout 200 python tools/prof_spmm.py > gpurun_out/r2l_spmm_plain.log 2>&1 && tail -n 1 gpurun_out/r2l_spmm_plain.log &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_spmm" -s 3 -c 1 -o gpurun_out/r2l_spmm python tools/prof_spmm.py > gpurun_out/r2l_spmm_ncu.log 2>&1; tail -n 3 gpurun_out/r2l_spmm_ncu.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest_gpu.log 2>&1; tail -n 8 gpurun_out/r2l_pytest_gpu.log
