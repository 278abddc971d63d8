# A/B of library variants: bench line summary per variant
mkdir -p gpurun_out
for v in "$@"; do
  export REGT_B200_LIB=$PWD/regt-gcn_b200/lib/variants/$v/libregt_b200.so
  timeout 120 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || { echo "$v FAILED"; tail -3 gpurun_out/ab_$v.err; continue; }
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open(f"gpurun_out/ab_{v}.json").read().strip().splitlines()[-1])
k=d["kernels"]
print(f"{v:18s} step {d['ms_per_step']*1e3:6.1f}us e2e {d['e2e']['ms_per_step']*1e3:6.1f}us | " + " ".join(f"{n.replace('k_','')}={x['ms_per_step']*1e3:.1f}" for n,x in k.items()))
PY
done
