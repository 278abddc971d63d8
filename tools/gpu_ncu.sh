# round evidence: ncu launch lists + --set full captures (bf16 default workload; tf32x3 config 3 GEMMs)
set -x
mkdir -p gpurun_out
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ncu2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_cell_(fwd|bwd)_tc|k_head_tc|k_feat_tc' -s 24 -c 4 -f -o gpurun_out/prof_v8_bf16 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
timeout 200 python bench.py --precision tf32x3 --workload 3 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_ncu3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v8_cfg3_tf32x3.csv python bench.py --precision tf32x3 --workload 3 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_gemm_(nt|tn)_tma|k_g_zr|k_g_b1' -s 16 -c 8 -f -o gpurun_out/prof_v8_cfg3_tf32x3 python bench.py --precision tf32x3 --workload 3 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full3.log 2>&1
tail -n 2 gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep
