# round 2, call AO: is the late L2 prefetch of the backward still worth its DRAM traffic at config 5?  (the kernel now runs at 4.5 TB/s)
set -x
mkdir -p gpurun_out
for v in base prefetch0 base prefetch0; do
  if [ $v = base ]; then unset REGT_B200_LIB; else export REGT_B200_LIB=regt-gcn_b200/lib/variants/$v/libregt_b200.so; fi
  timeout 400 python bench.py --workload 5 --no-extras --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/r2ao_b5_$v.json 2> gpurun_out/r2ao_b5_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2ao_b5_$v.json').read().strip().splitlines()[-1])
k=d['kernels']
print('$v', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
done
