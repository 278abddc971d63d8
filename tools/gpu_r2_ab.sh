# round 2, call AB: backward: late L2 prefetch (at bar_1) vs none; forward with h parked in TMEM
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2ab_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2ab_fused_all.log | tail -n 5
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2ab_fused_all.log; exit 1; fi
for v in base prefetch0 base; do
  if [ $v = base ]; then unset REGT_B200_LIB; else export REGT_B200_LIB=regt-gcn_b200/lib/variants/$v/libregt_b200.so; fi
  timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2ab_b4_$v.json 2> gpurun_out/r2ab_b4_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2ab_b4_$v.json').read().strip().splitlines()[-1])
k=d['kernels']
print('$v', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')})
PY
done
unset REGT_B200_LIB
timeout 200 python tools/f_phases.py > gpurun_out/r2ab_phases.log 2>&1; grep -A13 "backward" gpurun_out/r2ab_phases.log | head -15
