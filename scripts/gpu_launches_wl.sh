#!/bin/bash
# launch list of one secondary workload (kernels launched from the host, no CUDA graph): WL=config3|config4|config5
mkdir -p gpurun_out
WL=${WL:-config3}
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/plain_$WL.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-200} --csv --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_$WL.log 2>&1
echo "ncu rc=$?"
python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_$WL.csv')))
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
    tot[r[kn][:110]][0] += 1
    tot[r[kn][:110]][1] += v
s = sum(v for _, v in tot.values())
with open('gpurun_out/launch_share_$WL.txt', 'w') as f:
    f.write('## $WL: ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-200}: $CMD\n')
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        line = f'{100 * v / s:6.2f}%  n={n:4d}  avg={v / n:9.2f} us  {k}'
        print(line); f.write(line + '\n')
PY
