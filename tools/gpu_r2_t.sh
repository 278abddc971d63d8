# round 2, call T: do the SMs' store phases collide?  (fused kernels with staggered CTA starts)
set -x
mkdir -p gpurun_out
for s in 0 2000 4000 7000; do
REGT_F_STAGGER=$s timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2t_b4_s$s.json 2> gpurun_out/r2t_b4_s$s.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_b4_s$s.json').read().strip().splitlines()[-1])
k=d['kernels']
print('stagger $s', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')})
PY
done
REGT_F_STAGGER=4000 timeout 200 python tools/f_phases.py > gpurun_out/r2t_phases_s4000.log 2>&1; grep -A12 "backward" gpurun_out/r2t_phases_s4000.log | head -16
