# round 2, call AP: k_gemm_kt with advisory pacing of the M-tile CTAs of a split (shared h^T / F^T boxes stay in L2)
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2ap_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 30 gpurun_out/r2ap_fused_all.log; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_large_configs.py -m gpu -q -x > gpurun_out/r2ap_pytest.log 2>&1; tail -n 2 gpurun_out/r2ap_pytest.log
for pace in 1 0 1; do
  REGT_KT_PACE=$pace timeout 400 python bench.py --workload 5 --no-extras --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/r2ap_b5_$pace.json 2> gpurun_out/r2ap_b5_$pace.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2ap_b5_$pace.json').read().strip().splitlines()[-1])
k=d['kernels']
print('pace $pace', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
done
