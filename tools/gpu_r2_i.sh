# round 2, call I: where E0 of the fused backward spends its 18 us (finer marks, two timing experiments), spmm with 4 edges in flight
set -x
mkdir -p gpurun_out
timeout 200 python tools/f_phases.py > gpurun_out/r2i_phases.log 2>&1; grep -A16 "backward" gpurun_out/r2i_phases.log
for v in xp_noh16 xp_nostore; do
  REGT_B200_LIB=regt-gcn_b200/lib/variants/$v/libregt_b200.so timeout 200 python tools/f_phases.py > gpurun_out/r2i_phases_$v.log 2>&1; echo "== $v"; grep -A16 "backward" gpurun_out/r2i_phases_$v.log
done
timeout 300 python - <<'PY' > gpurun_out/r2i_spmm.log 2>&1
import sys, os, json
sys.path.insert(0, "."); sys.path.insert(0, "regt-gcn_b200")
import torch, bench
from regt_b200 import workloads as W
w = W.make_workload(5)
dev = torch.device("cuda:0")
x, _ = w.inputs(64)
g = tuple(None if a is None else a.to(dev) for a in w.graph_args())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(json.dumps(bench.measure_spmm(w, dev, x.to(dev), g, flush, bench.peaks())))
os.environ["REGT_SPMM_PLAIN"] = "1"
PY
tail -n 2 gpurun_out/r2i_spmm.log | cut -c1-400
