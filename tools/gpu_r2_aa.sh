# round 2, call AA: forward: next-step h parked in TMEM (fewer spills)
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2aa_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2aa_fused_all.log | tail -n 5
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2aa_fused_all.log; exit 1; fi
timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2aa_b4.json 2> gpurun_out/r2aa_b4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aa_b4.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg4', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')})
PY
timeout 200 python tools/f_phases.py > gpurun_out/r2aa_phases.log 2>&1; grep -B12 "backward" gpurun_out/r2aa_phases.log | head -14
