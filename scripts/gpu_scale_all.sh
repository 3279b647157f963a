#!/bin/bash
# 1 -> 8 GPU weak scaling of the default workload (config 2) and of config 5 on ONE 8-GPU box
set -u
mkdir -p gpurun_out
: > gpurun_out/scale_all.jsonl
run() { # N workload-args...
  N=$1; shift
  if [ $N -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary "$@" 2> gpurun_out/scale.err | tail -1 >> gpurun_out/scale_all.jsonl
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-secondary "$@" 2> gpurun_out/scale.err | tail -1 >> gpurun_out/scale_all.jsonl; fi
  echo "N=$N $* rc=$?"
}
for N in 1 2 4 8; do run $N; done
for N in 1 8; do run $N --workload config5; done
python - <<'PY'
import json
for l in open('gpurun_out/scale_all.jsonl'):
    try: j = json.loads(l)
    except Exception: print('bad line', l[:100]); continue
    print(j['config']['workload'][:40], 'N=%d' % j['n_gpus'], 'value %.2f M' % (j['value'] / 1e6), 'e2e %.2f M' % (j['e2e']['value'] / 1e6), '%.2f ms' % j['ms_per_step'], j['clocks']['sm_mhz'])
PY
