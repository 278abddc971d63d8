# round 2, call AF: how deep does the forward's weight ring have to be?  (3 / 4 / 5 stages)
set -x
mkdir -p gpurun_out
for v in base fns4 fns3 base; do
  if [ $v = base ]; then unset REGT_B200_LIB; else export REGT_B200_LIB=regt-gcn_b200/lib/variants/$v/libregt_b200.so; fi
  timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2af_b4_$v.json 2> gpurun_out/r2af_b4_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2af_b4_$v.json').read().strip().splitlines()[-1])
k=d['kernels']
print('$v', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
done
