#!/bin/bash
# One GPU-box visit: GPU tests, the bench line of the headline config, A/B of the lane dealing, ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/ab_lane_map.py 2>&1 | tail -5 | cut -c1-900 | tee gpurun_out/ab_lane_map.jsonl
python bench.py --config 2 --steps 1000 --warmup 20 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
python - <<PY
import json
l=json.load(open('gpurun_out/bench_c2.json'))
print('config 2', round(l['value']), 'ms', round(l['ms_per_step'],4), 'render', l['roofline']['ms_per_launch'], 'frac', round(l['roofline']['frac'],3), 'whole', l['roofline']['whole_step_frac'], 'e2e', l.get('e2e',{}).get('value'))
PY
ncu --set full --clock-control none --import-source on -k regex:k_render -s 4 -c 1 -f -o gpurun_out/render_r02 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/ncu_render.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -f -o gpurun_out/step_r02 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/ncu_step.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out | tail -8
