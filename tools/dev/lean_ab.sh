#!/bin/bash
out=gpurun_out/lean
mkdir -p $out
python -m pytest tests/test_map_gpu.py tests/test_views_gpu.py tests/test_policy_device_gpu.py tests/test_collect_gpu.py -x -q -m gpu > $out/tests.log 2>&1; echo "rc=$?" >> $out/tests.log
tail -2 $out/tests.log
for lean in 0 1; do
  MG_MAP_LEAN=$lean python tools/kbench_families.py --which ctf,ctf_policy,maze_partial > $out/fam_lean$lean.jsonl 2> $out/fam_lean$lean.err
done
python - <<'P'
import json
for lean in (0,1):
    for l in open(f'gpurun_out/lean/fam_lean{lean}.jsonl'):
        try: d=json.loads(l)
        except: continue
        print(lean, d['kernel'][:70], d['num_envs'], round(d['us_per_launch'],2), round(d['frac_of_measured_peak'],3))
P
