# round 2: the default bench line on the final commit (spmm key timed at the kernel)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2y_b5.json 2> gpurun_out/r2y_b5.err; tail -n 1 gpurun_out/r2y_b5.err; cut -c1-200 gpurun_out/r2y_b5.json
