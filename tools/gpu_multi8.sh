#!/bin/bash
# 8-GPU visit: bench line of every BASELINE config under torchrun (config 2 with the observation gather too).
N=${1:-8}
mkdir -p gpurun_out
run() {  # config steps extra
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $1 --steps $2 --warmup 20 $3 > gpurun_out/bench_c$1_${N}gpu.json 2> gpurun_out/bench_c$1_${N}gpu.err || tail -20 gpurun_out/bench_c$1_${N}gpu.err
  python - <<PY
import json
try:
  l=json.load(open('gpurun_out/bench_c$1_${N}gpu.json'))
  g=l.get('gather_obs') or {}
  print('config $1 N=$N', round(l['value']), 'ms', round(l['ms_per_step'],4), 'shard_check', l.get('shard_check'), 'e2e', l.get('e2e',{}).get('value'), 'frac_pcie', l.get('e2e',{}).get('frac_of_pcie'), 'gather', g.get('value'), g.get('nvlink_gbs_per_gpu_egress'), g.get('check'))
  for j in l['per_substrate']: print('   ', j['substrate'], j['players'], round(j['env_steps_per_sec']), round(j['render_frac'],3), round(j['whole_step_frac'],3))
except Exception as e: print('config $1 failed', e)
PY
}
run 2 1000 --gather-obs
run 4 500 --gather-obs
run 3 300 ""
run 5 300 ""
