#!/bin/bash
# Steady-state DRAM bytes per launch of every kernel family: application replay, no cache flush, three consecutive launches.
out=${1:-gpurun_out/ncu_traffic_families.txt}
: > $out
declare -A K=([ctf]=map_kernel [maze]=map_kernel [view_maze]=view_ [view_collect]=view_ [toroid]=toroid_ [wildfire]=wildfire_ [generic]=generic_kernel [render]=render_kernel)
for f in ctf maze view_maze view_collect toroid wildfire generic render; do
  echo "### $f" >> $out
  timeout 200 ncu --replay-mode application --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:${K[$f]} -s 6 -c 3 \
      python tools/profile_families.py $f 2>&1 | grep -v "^==PROF==" >> $out
done
