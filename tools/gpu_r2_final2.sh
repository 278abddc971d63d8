# round 2, final evidence 2: ncu --set full of the top kernels of the default bench command (config 5, tf32x3, one B200)
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-graph > gpurun_out/r2g_plain.json 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_cell_.*_f|k_gemm_kt|k_wgrad_m1_kt|k_feat_tc|k_hf_mid" -s 24 -c 6 -f -o gpurun_out/r02_cfg5_tf32x3_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-graph > gpurun_out/r2g_ncu_full.log 2>&1
tail -n 3 gpurun_out/r2g_ncu_full.log; ls -la gpurun_out/r02_cfg5_tf32x3_full.ncu-rep
