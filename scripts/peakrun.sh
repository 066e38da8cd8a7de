mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > gpurun_out/peak_clocks.csv &
SMI=$!
for b in 1 2 3 4 6; do echo "== blocks/SM=$b"; NLS_PEAK_BLOCKS_PER_SM=$b python - <<'PY'
import sys; sys.path.insert(0,'.')
from neo_ls_svm_b200 import _lib
ctx=_lib.Context(0)
print([round(ctx.dmma_peak_tflops(20000),2) for _ in range(3)], [round(ctx.dmma_peak_tflops(200000),2) for _ in range(2)])
PY
done
kill $SMI
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/peak_clocks.csv'))][1:]
clk=[int(r[0].split()[0]) for r in rows if r]; pw=[float(r[1].split()[0]) for r in rows if r]
busy=[(c,p,r[2]) for c,p,r in zip(clk,pw,rows) if p>400]
print('samples',len(rows),'busy',len(busy),'min/max clk busy',min(c for c,_,_ in busy) if busy else None,max(c for c,_,_ in busy) if busy else None,'max power',max(pw), 'power_cap active', sum(1 for _,_,a in busy if 'Active' in a and 'Not' not in a))
PY
