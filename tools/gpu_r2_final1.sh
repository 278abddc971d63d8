# round 2, final evidence 1: fused check, the whole GPU suite, the default bench (config 5, with extras) and configs 4 / 3 / 2,
# ncu launch list of the default command
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2f_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2f_fused_all.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -n 6 gpurun_out/r2f_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -n 3 gpurun_out/r2f_smoke.log
timeout 600 python bench.py > gpurun_out/r2f_b5.json 2> gpurun_out/r2f_b5.err; tail -n 2 gpurun_out/r2f_b5.err; cut -c1-200 gpurun_out/r2f_b5.json
timeout 400 python bench.py --workload 4 --no-extras --no-cpu-baseline > gpurun_out/r2f_b4.json 2> gpurun_out/r2f_b4.err; cut -c1-200 gpurun_out/r2f_b4.json
timeout 400 python bench.py --workload 3 --no-extras --no-cpu-baseline > gpurun_out/r2f_b3.json 2> gpurun_out/r2f_b3.err; cut -c1-200 gpurun_out/r2f_b3.json
timeout 400 python bench.py --workload 2 --no-extras --no-cpu-baseline > gpurun_out/r2f_b2.json 2> gpurun_out/r2f_b2.err; cut -c1-200 gpurun_out/r2f_b2.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-graph > gpurun_out/r2f_plain.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg5_tf32x3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-graph > gpurun_out/r2f_ncu_launches.log 2>&1
tail -n 2 gpurun_out/r2f_ncu_launches.log; wc -l gpurun_out/r02_launches_cfg5_tf32x3.csv
