#!/bin/bash
# Round-2 closing pass (after the lean map kernels and the pooled host decode): GPU tests, bench lines, launch list.
out=gpurun_out/r02c
mkdir -p $out
t0=$(date +%s)
python -m pytest tests -x -q -m gpu > $out/gputests.log 2>&1; echo "rc=$?" >> $out/gputests.log
tail -3 $out/gputests.log; echo "tests: $(( $(date +%s) - t0 )) s"
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err || { echo "bench failed"; tail -5 $out/bench_n1.err; }
echo "bench: $(( $(date +%s) - t0 )) s"
python bench.py --steps 20 --warmup 5 > $out/bench_n1_driver_flags.json 2>> $out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference_arm.json 2>> $out/bench_n1.err
echo "arms: $(( $(date +%s) - t0 )) s"
python bench.py --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 4 --e2e-repeats 1 --skip-families --skip-rollout --skip-small --repeats 2 > $out/plain_bench_short.json 2>> $out/bench_n1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv \
    python bench.py --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 4 --e2e-repeats 1 --skip-families --skip-rollout --skip-small --repeats 2 > $out/ncu_bench.log 2>&1
python tools/kbench_families.py --which ctf,ctf_policy,maze,maze_partial,view,wildfire,generic > $out/kbench_families.jsonl 2> $out/kbench_families.err
echo "total: $(( $(date +%s) - t0 )) s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02c/bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
for k,v in d.get('families',{}).items(): print(k, v.get('ms_per_step'), v.get('roofline',{}).get('frac'))
P
