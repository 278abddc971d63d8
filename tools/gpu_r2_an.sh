# round 2, call AN: forward: h of the step kept in registers (no TMEM re-reads of the A operand in E1z / E1r)
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2an_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"; grep -v "grad " gpurun_out/r2an_fused_all.log | tail -n 6; awk '/^\[/{c=$1} /grad/{ if ($4+0 > 4e-6) print c, $2, $4}' gpurun_out/r2an_fused_all.log | head
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 30 gpurun_out/r2an_fused_all.log; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_large_configs.py tests/test_gpu_model_parity.py -m gpu -q -x > gpurun_out/r2an_pytest.log 2>&1; tail -n 4 gpurun_out/r2an_pytest.log
for i in 1 2; do
timeout 300 python bench.py --workload 4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2an_b4.json 2> gpurun_out/r2an_b4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2an_b4.json').read().strip().splitlines()[-1])
k=d['kernels']
print('cfg4', d['ms_per_step'], {n:k[n]['ms_per_step'] for n in ('k_cell_bwd_f','k_cell_fwd_f','k_gemm_kt')}, d['clocks']['sm_mhz'])
PY
done
timeout 200 python tools/f_phases.py > gpurun_out/r2an_phases.log 2>&1; grep -B11 "backward" gpurun_out/r2an_phases.log | head -12; grep -A9 "MMA warp" gpurun_out/r2an_phases.log
