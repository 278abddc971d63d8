# round 2, call V: is E0 of the backward kernel bound per SM or by the chip (DRAM)?  phase times with 148 / 74 / 37 CTAs
set -x
mkdir -p gpurun_out
for n in 148 74 37; do
REGT_F_SMS=$n timeout 200 python tools/f_phases.py > gpurun_out/r2v_phases_$n.log 2>&1; echo "CTAs $n"; grep -B11 -A14 "backward" gpurun_out/r2v_phases_$n.log | grep "E1z\|E1r\|E2 \|E0\|wait M\|step to step"
done
