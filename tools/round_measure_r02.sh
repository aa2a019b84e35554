#!/bin/bash
# Round-2 measurement pass for profiles/ (run under gpurun; everything lands in gpurun_out/r02m/):
#   bench lines (default K, the driver's flags, the reference arm), the ncu launch list of the bench command, the Collect variants,
#   rollouts, the other families, one ncu --set full capture of the Collect step (warp-tile kernel, T = 1) with its source page, and
#   the steady-state DRAM bytes per launch.
out=gpurun_out/r02m
mkdir -p $out
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err || { echo "bench failed"; tail -5 $out/bench_n1.err; }
python bench.py --steps 20 --warmup 5 > $out/bench_n1_driver_flags.json 2>> $out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference_arm.json 2>> $out/bench_n1.err
python bench.py --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 4 --e2e-repeats 1 --skip-families --skip-rollout --skip-small --repeats 2 > $out/plain_bench_short.json 2>> $out/bench_n1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv \
    python bench.py --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 4 --e2e-repeats 1 --skip-families --skip-rollout --skip-small --repeats 2 > $out/ncu_bench.log 2>&1
python tools/kbench_collect.py --constant --num-envs 65536,16384,4096 --variants "MG_STEP_IMPL=warp;MG_STEP_IMPL=tile,MG_TILE=9;MG_STEP_IMPL=tile,MG_TILE=0;MG_STEP_IMPL=tile,MG_TILE=0,MG_EARLY_OBS=1" > $out/kbench_collect_variants.jsonl 2> $out/kbench_collect.err
python tools/kbench_rollout.py --num-envs 4096,16384,65536 --steps 1,8,64 --modes actions,policy,noobs > $out/kbench_rollout.jsonl 2> $out/kbench_rollout.err
python tools/kbench_families.py --which ctf,ctf_policy,maze,maze_partial,view,wildfire,generic,render > $out/kbench_families.jsonl 2> $out/kbench_families.err
python tools/profile_collect.py > $out/plain_collect.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"collect_rollout|collect_step" -s 24 -c 1 -f -o $out/prof_collect_step python tools/profile_collect.py > $out/ncu_collect.log 2>&1
timeout 200 ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -k regex:"collect_rollout|collect_step" -s 24 -c 8 python tools/profile_collect.py 2>&1 | grep -v "^==PROF==" > $out/ncu_collect_steady_traffic.txt
echo "steady traffic: $(grep -c dram__bytes_read $out/ncu_collect_steady_traffic.txt) launches"
