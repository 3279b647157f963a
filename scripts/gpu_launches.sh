#!/bin/bash
# launch list of the default bench command (CUDA-graph replay), shares of summed kernel time
mkdir -p gpurun_out
CMD="python bench.py --k-step 40 --steps 2 --warmup 3 --no-cpu-baseline ${BARGS:-}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-600} --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_r02.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
kn, mv = rows[hdr].index('Kernel Name'), rows[hdr].index('Metric Value')
mu = rows[hdr].index('Metric Unit')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(',', ''))
    if r[mu] == 'ns': v /= 1e3
    elif r[mu] == 'ms': v *= 1e3
    elif r[mu] in ('s', 'second'): v *= 1e6
    tot[r[kn][:100]][0] += 1
    tot[r[kn][:100]][1] += v
s = sum(v for _, v in tot.values())
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{100 * v / s:6.2f}%  n={n:4d}  avg={v / n:9.2f} us  {k}')
PY
