#!/bin/bash
# ncu --set full capture of one launch per kernel family -> gpurun_out/prof_<family>_<tag>.ncu-rep  (run under gpurun)
tag=${1:-r1}
shift
fams=${@:-"ctf view_maze view_collect toroid wildfire generic render"}
declare -A K=([ctf]=map_kernel [ctf8]=map_kernel [maze_partial]=map_kernel [ctf_policy]=ctf_policy_kernel [maze]=map_kernel [view_maze]=view_ [view_collect]=view_ [toroid]=toroid_ [wildfire]=wildfire_ [generic]=generic_kernel [render]=render_kernel)
for f in $fams; do
  python tools/profile_families.py $f > gpurun_out/plain_$f.log 2>&1 || { echo "plain run of $f failed"; tail -5 gpurun_out/plain_$f.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:${K[$f]} -s 6 -c 1 -f -o gpurun_out/prof_${f}_${tag} \
      python tools/profile_families.py $f > gpurun_out/ncu_$f.log 2>&1
  echo "$f: rc=$?"
done
