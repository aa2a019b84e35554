#!/bin/bash
# One measurement pass for profiles/: bench line, ncu launch list of the same command, per-family micro-benchmarks,
# one ncu --set full capture per kernel family.  Run under gpurun; everything lands in gpurun_out/.
tag=${1:-r01}
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err || { echo "bench failed"; tail -5 gpurun_out/${tag}_bench_n1.err; }
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 4 > gpurun_out/${tag}_ncu_bench.log 2>&1
python tools/kbench_families.py --which memset,ctf,maze,maze_partial,view,wildfire,generic,render,collect_streams > gpurun_out/${tag}_kbench_families.jsonl 2> gpurun_out/${tag}_kbench_families.err
python tools/kbench.py --tiles 0 --num-envs 65536 > gpurun_out/${tag}_kbench_collect_65536.log 2>&1
python tools/kbench.py --tiles 0 --num-envs 1048576 --batches 4 --steps 400 > gpurun_out/${tag}_kbench_collect_1M.log 2>&1
bash tools/ncu_families.sh ${tag} ctf maze view_maze view_collect toroid wildfire generic render
python tools/profile_collect.py > gpurun_out/plain_collect.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:collect_step -s 20 -c 1 -f -o gpurun_out/prof_collect_${tag} python tools/profile_collect.py > gpurun_out/ncu_collect.log 2>&1
echo "collect: rc=$?"
# steady-state DRAM bytes per launch of the Collect step (roofline.traffic): application replay, no cache flush, launches 30-37 of a rotation
timeout 200 ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -k regex:collect_step -s 30 -c 8 python tools/profile_collect.py 2>&1 | grep -v "^==PROF==" > gpurun_out/${tag}_ncu_collect_steady_traffic.txt
echo "steady traffic: $(grep -c dram__bytes_read gpurun_out/${tag}_ncu_collect_steady_traffic.txt) launches"
