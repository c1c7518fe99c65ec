#!/bin/bash
# Final 1-GPU visit of the round: the whole GPU suite, smoke, the bench lines, ncu captures of the current kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
for c in 3 4 5; do python bench.py --config $c --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err || tail -5 gpurun_out/bench_c$c.err; done
python - <<PY
import json
for c in (2,3,4,5):
  try:
    l=json.load(open('gpurun_out/bench_c%d.json'%c))
    print('config',c, round(l['value']), 'ms', round(l['ms_per_step'],4), 'render', round(l['roofline']['ms_per_launch'],4), round(l['roofline']['frac'],3), 'whole', l['roofline']['whole_step_frac'], 'traffic', l['roofline']['traffic'], 'e2e', l.get('e2e',{}).get('value'), l.get('e2e',{}).get('frac_of_pcie'), 'cpu', l.get('cpu_baseline',{}).get('value'))
    for j in l['per_substrate']: print('   ', j['substrate'], j['players'], round(j['env_steps_per_sec']), round(j['render_frac'],3), round(j['whole_step_frac'],3))
  except Exception as e: print('config',c,'failed', e)
PY
python tools/throughput_all.py > gpurun_out/throughput_r02.jsonl 2> gpurun_out/throughput.err
ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -f -o gpurun_out/step_r02 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -f -o gpurun_out/step_territory_r02 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/ncu_step_t.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 4 > gpurun_out/ncu_launches.log 2>&1
