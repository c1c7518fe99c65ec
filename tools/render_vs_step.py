"""Render-kernel time as the episode progresses (diagnostic)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meltingpot_b200 import engine, substrates

def timeit(fn, n=20):
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / n

blob = substrates.load_blob('clean_up')
eng = engine.Engine(blob, 4096, seed=1)
eng.reset()
gen = torch.Generator(device='cuda').manual_seed(0)
t = 0
for target in (0, 50, 200, 600, 1000, 1500, 2000):
  while t < target:
    eng.step_state(torch.randint(0, 9, (4096, 7), generator=gen, device='cuda', dtype=torch.int32)); t += 1
  full = timeit(eng.render)
  eng.set_flags(3 | 16); stores = timeit(eng.render); eng.set_flags(3)
  alive = float(eng.avatar_state[:, :, 3].float().mean())
  grid = eng.grid.view(torch.int16)
  dirt = float((grid[:, 4] != 0).float().sum(1).mean())
  beams = float((grid[:, 7:9] != 0).float().sum((1, 2)).mean())
  print(json.dumps({'step': t, 'render_ms': full, 'stores_only_ms': stores, 'alive_frac': alive, 'upperPhysical_cells': dirt, 'beam_cells': beams}))
