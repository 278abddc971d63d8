# round 2, call AD: k_gemm_kt with work-weighted splits (the feature-only M tile gets fewer CTAs)
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2ad_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2ad_fused_all.log | tail -n 5
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2ad_fused_all.log; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py -m gpu -q -x > gpurun_out/r2ad_pytest.log 2>&1; tail -n 3 gpurun_out/r2ad_pytest.log
for i in 1 2; do
timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2ad_b4.json 2> gpurun_out/r2ad_b4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ad_b4.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg4', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
done
timeout 400 python bench.py --workload 5 --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2ad_b5.json 2> gpurun_out/r2ad_b5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ad_b5.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg5', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
