# round 2, call K: tensor-core head (head_f.cu), spmm v3 (pre-resolved edges), plain stores in the backward again
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2k_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2k_fused_all.log | tail -n 8; awk '/^\[/{c=$1} /grad/{ if ($4+0 > 4e-6) print c, $2, $4}' gpurun_out/r2k_fused_all.log
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2k_fused_all.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest_gpu.log 2>&1; tail -n 8 gpurun_out/r2k_pytest_gpu.log
timeout 400 python bench.py --workload 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_b5.json 2> gpurun_out/r2k_b5.err; tail -n 3 gpurun_out/r2k_b5.err; cut -c1-300 gpurun_out/r2k_b5.json
timeout 200 python tools/prof_spmm.py 2>&1 | tail -n 1
PB=64 timeout 200 python tools/prof_spmm.py 2>&1 | tail -n 1
timeout 300 python bench.py --workload 3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_b3.json 2> gpurun_out/r2k_b3.err; cut -c1-300 gpurun_out/r2k_b3.json
timeout 300 python bench.py --workload 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_b2.json 2> gpurun_out/r2k_b2.err; cut -c1-300 gpurun_out/r2k_b2.json
