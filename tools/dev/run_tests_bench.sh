#!/bin/bash
out=gpurun_out/r02b
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/gputests.log 2>&1; echo "rc=$?" >> $out/gputests.log
tail -3 $out/gputests.log
python bench.py --steps 2000 --warmup 16 --repeats 3 --skip-families --skip-rollout --skip-small > $out/bench_short.json 2> $out/bench_short.err || tail -5 $out/bench_short.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02b/bench_short.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value',d['value'],'e2e',e['value'],'packed',e['packed']['value'],'full',e['full']['value'],'blocking',e['blocking']['value'],'cpu',d.get('cpu_baseline',{}).get('value'))
P
