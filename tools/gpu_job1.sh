set -x
mkdir -p gpurun_out
timeout 300 python bench.py --precision bf16 > gpurun_out/b2_bf16.json 2> gpurun_out/b2_bf16.err
timeout 300 python bench.py --precision fp32 --workload 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b3_fp32.json 2> gpurun_out/b3_fp32.err
timeout 400 python bench.py --precision fp32 --workload 4 --micro-batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/b4_fp32.json 2> gpurun_out/b4_fp32.err
timeout 300 python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_cell_(fwd|bwd)_tc|k_head_fused|k_feat_tc' -s 24 -c 4 -f -o gpurun_out/prof_r01_v4 python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
