mkdir -p gpurun_out
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:60]:60s} n={len(v):4d} avg={sum(v)/len(v)/1000:8.2f} us  min={min(v)/1000:8.2f}")
PY
