# round 2, call Z: backward E0 as a rolled 8-column loop, Dz in shared memory (no spills), streaming cache operators
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2z_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2z_fused_all.log | tail -n 5
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 20 gpurun_out/r2z_fused_all.log; exit 1; fi
timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2z_b4.json 2> gpurun_out/r2z_b4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2z_b4.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg4', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')})
PY
timeout 200 python tools/f_phases.py > gpurun_out/r2z_phases.log 2>&1; grep -A13 "backward" gpurun_out/r2z_phases.log | head -15
timeout 400 python bench.py --workload 5 --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2z_b5.json 2> gpurun_out/r2z_b5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2z_b5.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg5', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')})
PY
timeout 200 python tools/prof_cell.py > gpurun_out/r2z_prof_plain.log 2>&1 && tail -n 1 gpurun_out/r2z_prof_plain.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell_bwd_f|k_cell_fwd_f" -s 2 -c 2 -f -o gpurun_out/r2z_cell python tools/prof_cell.py > gpurun_out/r2z_ncu.log 2>&1; tail -n 2 gpurun_out/r2z_ncu.log
