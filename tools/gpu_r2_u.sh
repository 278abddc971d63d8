# round 2, call U: store width microbenchmark; MN-major no-swizzle operand probe
set -x
mkdir -p gpurun_out
timeout 60 tools/_bin/store_bw > gpurun_out/r2u_store_bw.log 2>&1; cat gpurun_out/r2u_store_bw.log
timeout 100 python tools/umma_probe2.py > gpurun_out/r2u_umma.log 2>&1; cat gpurun_out/r2u_umma.log | tail -n 14
