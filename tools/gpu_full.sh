# round evidence run (1 GPU): GPU tests, smoke, default bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 6 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -n 2 gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json
