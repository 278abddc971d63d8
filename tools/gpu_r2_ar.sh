# round 2, call AR: dM1 weight gradient with 32 warps x 4 columns instead of 16 x 8
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2ar_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 30 gpurun_out/r2ar_fused_all.log; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_large_configs.py -m gpu -q -x > gpurun_out/r2ar_pytest.log 2>&1; tail -n 2 gpurun_out/r2ar_pytest.log
for wv in 0 1; do
REGT_M1_WIDE=$wv timeout 400 python bench.py --workload 5 --no-extras --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/r2ar_b5_$wv.json 2> gpurun_out/r2ar_b5_$wv.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ar_b5_$wv.json').read().strip().splitlines()[-1])
k=d['kernels']
print('wide=$wv', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt','k_wgrad_m1_kt')}, d['clocks']['sm_mhz'])
PY
done
