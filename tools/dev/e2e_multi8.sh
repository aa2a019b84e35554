#!/bin/bash
# eight processes, one per GPU, started together (the box's host cores are shared): e2e of the host-buffer path per rank
out=gpurun_out/e2e8
mkdir -p $out
{ nproc; lscpu | grep -i "model name\|socket\|numa\|thread\|L2\|L3"; lscpu -e | head -40; nvidia-smi topo -m; } > $out/topology.txt 2>&1
G=${1:-8}
run() {  # name, env assignments, configs
  name=$1; shift; envs=$1; shift
  for i in $(seq 0 $((G-1))); do
    env CUDA_VISIBLE_DEVICES=$i LOCAL_WORLD_SIZE=$G E2E_ALIGN=10 E2E_STEPS=300 E2E_REPS=2 $envs python tools/dev/e2e_async.py "$@" > $out/${name}_rank$i.log 2>&1 &
  done
  wait
  echo "== $name"; grep -h "EB=" $out/${name}_rank*.log | awk '{print $7, $8, $10}' | sort | uniq -c | head -0
  python - "$out" "$name" <<'P'
import glob, re, sys, collections
tot = collections.defaultdict(list)
for f in glob.glob(f"{sys.argv[1]}/{sys.argv[2]}_rank*.log"):
    for l in open(f):
        m = re.search(r"(EB=\d+ threads=\s*\d+): best ([0-9.e+]+)", l)
        if m: tot[m.group(1)].append(float(m.group(2)))
for k, v in tot.items():
    print(f"{sys.argv[2]} {k}: ranks={len(v)} sum={sum(v):.3e} min={min(v):.3e} max={max(v):.3e}")
P
}
run nopin "E2E_X=0" 1,2 1,3 1,4 2,4
run pin "E2E_PIN=1" 1,2 1,4 2,4
run pinht "E2E_PIN=ht" 1,4 2,4
