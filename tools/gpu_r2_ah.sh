# round 2, call AH: full GPU suite + configs 5 / 2 with the tensor-core h_pre in the forward
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2ah_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 30 gpurun_out/r2ah_fused_all.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2ah_pytest_gpu.log 2>&1; tail -n 4 gpurun_out/r2ah_pytest_gpu.log
for w in 5 2; do
timeout 400 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2ah_b$w.json 2> gpurun_out/r2ah_b$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ah_b$w.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg$w', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt') if n in k}, d['clocks']['sm_mhz'])
PY
done
