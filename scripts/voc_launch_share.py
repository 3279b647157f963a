"""Per-stage / per-kernel time of one vocoder call from an ncu launch list (scripts/gpu_vocoder_prof.sh)."""
import csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/voc_launches.csv'
lines = [l for l in open(path) if not l.startswith('==')]
rows = [(d['Kernel Name'], float(d['Metric Value'].replace(',', ''))) for d in csv.DictReader(lines) if d.get('Metric Name') == 'gpu__time_duration.sum']
own = [(k, t) for k, t in rows if 'b2s' in k or 'tc_gemm' in k]
# the script runs the call twice: keep the second
call = own[len(own) // 2:]
tot = sum(t for _, t in call)
print(f'{len(call)} launches per call, {tot / 1e3:.1f} us of kernel time (serialised under ncu)')
stage, agg = -1, {}
for k, t in call:
    name = re.sub(r'\(.*', '', k).replace('void ', '').replace('b2s::', '').replace('tc::', '')
    if 'voc_source_add' in name:
        stage += 1
    a = agg.setdefault(('pre' if stage < 0 else f'stage {stage}', name), [0, 0.0])
    a[0] += 1
    a[1] += t
for (st, name), (n, t) in agg.items():
    print(f'{st:8s} {name:36s} x{n:3d} {t / 1e3:8.1f} us  {100 * t / tot:5.1f} %')
