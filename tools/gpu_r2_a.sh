# round 2, call A: GPU tests (incl. the new large-config parity tests) + cfg5 / cfg4 tf32x3 bench lines with the chunked dM1 kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -n 15 gpurun_out/r2a_pytest_gpu.log
timeout 600 python bench.py --workload 5 --precision tf32x3 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2a_b5.json 2> gpurun_out/r2a_b5.err; tail -n 3 gpurun_out/r2a_b5.err; cut -c1-300 gpurun_out/r2a_b5.json
timeout 600 python bench.py --workload 4 --precision tf32x3 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2a_b4.json 2> gpurun_out/r2a_b4.err; tail -n 3 gpurun_out/r2a_b4.err; cut -c1-300 gpurun_out/r2a_b4.json
