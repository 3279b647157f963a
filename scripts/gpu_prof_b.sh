#!/bin/bash
# full capture of ONE kernel of the shipped path (KERNEL=regex) -> gpurun_out/r02_prof_$NAME.ncu-rep
set -u
mkdir -p gpurun_out
DBG="python bench.py --k-step 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-secondary ${BARGS:-}"
$DBG > gpurun_out/plain2.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-4} -c 1 -o gpurun_out/r02_prof_${NAME} -f $DBG > gpurun_out/ncu2.log 2>&1
echo "full ${NAME} rc=$?"
