# round 2, call AI: where does the forward's MMA warp spend the step?  (issue marks + weight-stage waits); h_pre path A/B test
set -x
mkdir -p gpurun_out
timeout 150 python tools/fused_check.py > gpurun_out/r2ai_fused_all.log 2>&1
rc=$?; echo "fused_check rc=$rc"
if [ $rc -ne 0 ]; then echo "FUSED CHECK FAILED: stopping"; tail -n 30 gpurun_out/r2ai_fused_all.log; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "hpre" > gpurun_out/r2ai_pytest.log 2>&1; tail -n 3 gpurun_out/r2ai_pytest.log
timeout 200 python tools/f_phases.py > gpurun_out/r2ai_phases.log 2>&1; grep -B11 "backward" gpurun_out/r2ai_phases.log | head -12; grep -A10 "MMA warp" gpurun_out/r2ai_phases.log
