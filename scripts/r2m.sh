cd "$(dirname "$0")/.."
python scripts/c2_trace.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_c2_launches.csv python scripts/c2_trace.py > /dev/null 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2m_c2_launches.csv")) if len(r) > 10 and r[0].isdigit()]
# last fit of the ozaki core = launches before the first launch of the dmma-sweep core; take launches by name totals for the first third
names = [r[4] for r in rows]
first_dmma = next(i for i, n in enumerate(names) if "OpSweep" in n)
seg = rows[:first_dmma]
third = seg[2 * len(seg) // 3:]
agg = collections.OrderedDict()
for r in third:
    a = agg.setdefault(r[4][:80], [0, 0.0]); a[0] += 1; a[1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("ozaki fit (last of 3): total kernel time %.2f ms in %d launches" % (tot / 1e6, len(third)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{v[1]/1e6:8.3f} ms {v[0]:5d} x  {k}")
P
