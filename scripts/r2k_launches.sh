#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2k_launches_bench_262k.csv python bench.py --steps 1 --warmup 1 --rows 262144 --skip-api --skip-configs > gpurun_out/r2k_ncu_bench.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2k_launches_bench_262k.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name, val = r[4], float(r[-1].replace(",", ""))
    a = agg.setdefault(name[:90], [0, 0.0])
    a[0] += 1; a[1] += val
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{v[1]/1e6:9.3f} ms {v[0]:5d} x {v[1]/v[0]/1e3:9.1f} us  {100*v[1]/tot:5.1f}%  {k}")
P
