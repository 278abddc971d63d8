# round 2, last call: the whole GPU suite, smoke(), the default bench on the final commit
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2h_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 30 gpurun_out/r2h_fused_all.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2z_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -n 1 gpurun_out/r2z_smoke.log
timeout 600 python bench.py > gpurun_out/r2z_b5.json 2> gpurun_out/r2z_b5.err; tail -n 1 gpurun_out/r2z_b5.err; cut -c1-220 gpurun_out/r2z_b5.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; cut -c1-400 gpurun_out/r2z_ref.json
