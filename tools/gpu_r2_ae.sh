# round 2, call AE: k_gemm_kt with an L2 prefetch of the chunk 4 / 8 ahead
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/r2ae_pytest.log 2>&1; tail -n 2 gpurun_out/r2ae_pytest.log
for v in base ktpf0 ktpf8 base; do
  if [ $v = base ]; then unset REGT_B200_LIB; else export REGT_B200_LIB=regt-gcn_b200/lib/variants/$v/libregt_b200.so; fi
  timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2ae_b4_$v.json 2> gpurun_out/r2ae_b4_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2ae_b4_$v.json').read().strip().splitlines()[-1])
k=d['kernels']
print('$v', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
done
unset REGT_B200_LIB
timeout 400 python bench.py --workload 5 --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2ae_b5.json 2> gpurun_out/r2ae_b5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ae_b5.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg5', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
