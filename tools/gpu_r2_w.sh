# round 2, call W: ncu --set full of the fused cell kernels + weight-gradient contraction (config 4, B = 16), current HEAD
set -x
mkdir -p gpurun_out
timeout 200 python tools/prof_cell.py > gpurun_out/r2w_prof_plain.log 2>&1 && tail -n 2 gpurun_out/r2w_prof_plain.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell_.*_f|k_gemm_kt|k_wgrad_m1_kt|k_feat_tc|k_hf_mid" -s 6 -c 6 -f -o gpurun_out/r2w_cell_f python tools/prof_cell.py > gpurun_out/r2w_ncu.log 2>&1; tail -n 3 gpurun_out/r2w_ncu.log
