# round 2, last check of the final tree: GPU suite + smoke
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest_gpu.log 2>&1; tail -n 2 gpurun_out/r2x_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; tail -n 1 gpurun_out/r2x_smoke.log
