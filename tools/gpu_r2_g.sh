# round 2, call G: phase timing of the fused kernels, L2 prefetch of the next step's planes, spmm order fix
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2g_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2g_fused_all.log | tail -n 8; awk '/^\[/{c=$1} /grad/{ if ($4+0 > 4e-6) print c, $2, $4}' gpurun_out/r2g_fused_all.log
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2g_fused_all.log; exit 1; fi
timeout 200 python tools/f_phases.py > gpurun_out/r2g_phases.log 2>&1; cat gpurun_out/r2g_phases.log | tail -n 30
timeout 400 python bench.py --workload 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_b5.json 2> gpurun_out/r2g_b5.err; tail -n 3 gpurun_out/r2g_b5.err; cut -c1-300 gpurun_out/r2g_b5.json
timeout 300 python bench.py --workload 4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_b4.json 2> gpurun_out/r2g_b4.err; cut -c1-300 gpurun_out/r2g_b4.json
