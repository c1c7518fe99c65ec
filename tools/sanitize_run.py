"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck) on both kernels.

  compute-sanitizer --tool memcheck  python tools/sanitize_run.py
  compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meltingpot_b200 import engine, substrates

for name, roles, n_act in (('clean_up', ('default',) * 7, 9), ('commons_harvest__open', ('default',) * 16, 8)):
  blob = substrates.load_blob(name, roles)
  eng = engine.Engine(blob, 300, seed=3)   # > 2 * 148 so that render teams process two envs each
  eng.reset()
  gen = torch.Generator(device='cuda').manual_seed(0)
  for _ in range(12):
    eng.step(torch.randint(0, n_act, (300, len(roles)), generator=gen, device='cuda', dtype=torch.int32))
  torch.cuda.synchronize()
  print(name, 'ok', int(eng.rgb.sum()) % 1000, eng.launch_count())
  eng.close()
