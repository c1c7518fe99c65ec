#!/bin/bash
# N-GPU visit: the 2-GPU tests, then the bench line under torchrun at N (default 2).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for c in ${CONFIGS:-2}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps ${STEPS:-1000} --warmup 20 ${EXTRA:-} > gpurun_out/bench_c${c}_${N}gpu.json 2> gpurun_out/bench_c${c}_${N}gpu.err || tail -20 gpurun_out/bench_c${c}_${N}gpu.err
python - <<PY
import json
try:
  l=json.load(open('gpurun_out/bench_c${c}_${N}gpu.json'))
  print('config $c N=$N', round(l['value']), 'ms', round(l['ms_per_step'],4), 'shard_check', l.get('shard_check'), 'e2e', l.get('e2e',{}).get('value'), 'frac_pcie', l.get('e2e',{}).get('frac_of_pcie'), 'numa', l.get('e2e',{}).get('numa'), 'launches', l['gpu_launches'], 'gather', l.get('gather_obs'))
except Exception as e: print('failed', e)
PY
done
