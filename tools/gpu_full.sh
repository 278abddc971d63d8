# round evidence run: GPU tests, smoke, default bench, ncu launch list, ncu --set full of the tensor-core kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ncu2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_cell_(fwd|bwd)_tc|k_head_tc|k_feat_tc' -s 24 -c 4 -f -o gpurun_out/prof_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
