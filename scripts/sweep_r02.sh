#!/bin/bash
# X6: batch x utterance-length sweep of configs 2 and 5 (one B200) + one ragged run.  Results: gpurun_out/sweep_r02.jsonl
set -u
mkdir -p gpurun_out
out=gpurun_out/sweep_r02.jsonl
: > $out
for wl in config2 config5; do
  for T in 345 1292 2584; do
    for B in 1 8 64 256; do
      ks=""; [ $wl = config2 ] && ks="--k-step 40"
      timeout 600 python bench.py --workload $wl --batch $B --frames $T $ks --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/sweep_tmp.log 2> gpurun_out/sweep_tmp.err
      rc=$?
      if [ $rc -eq 0 ]; then tail -1 gpurun_out/sweep_tmp.log | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print(json.dumps(dict(workload='$wl',B=$B,T=$T,value=j['value'],ms_per_step=j['ms_per_step'],e2e=j['e2e']['value'],frac=j['roofline']['frac'],kernel=j['roofline']['kernel'][:40],clocks=j['clocks'])))" >> $out
      else echo "{\"workload\":\"$wl\",\"B\":$B,\"T\":$T,\"rc\":$rc,\"err\":\"$(tail -1 gpurun_out/sweep_tmp.err | cut -c1-200 | tr -d '\"')\"}" >> $out; fi
      tail -1 $out | cut -c1-220
    done
  done
done
timeout 900 python scripts/ragged_bench.py > gpurun_out/ragged_r02.json 2> gpurun_out/ragged_r02.err; echo "ragged rc=$?"; cat gpurun_out/ragged_r02.json; tail -3 gpurun_out/ragged_r02.err
