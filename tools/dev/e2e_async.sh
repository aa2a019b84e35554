#!/bin/bash
out=gpurun_out/e2e
mkdir -p $out
lscpu | grep -i "L2\|L3" > $out/e2e_async.log
{
for m in 0 2 1; do
MG_DECODE_PREFETCH=$m E2E_REPS=2 python tools/dev/e2e_async.py 3,16 8,16
MG_DECODE_PREFETCH=$m E2E_REPS=2 taskset -c 0-3 python tools/dev/e2e_async.py 1,4 2,4 8,4
done
MG_DECODE_PREFETCH=2 MG_DECODE_PREFETCH_DIST=6 E2E_REPS=2 taskset -c 0-3 python tools/dev/e2e_async.py 8,4
MG_DECODE_PREFETCH=2 MG_DECODE_PREFETCH_DIST=24 E2E_REPS=2 taskset -c 0-3 python tools/dev/e2e_async.py 8,4
} >> $out/e2e_async.log 2>&1
cat $out/e2e_async.log
