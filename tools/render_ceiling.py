"""Render kernel vs its stores-only variant for one substrate (diagnostic)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meltingpot_b200 import engine, substrates

name, players, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])

def timeit(fn, n=20):
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / n

blob = substrates.load_blob(name, ('default',) * players)
eng = engine.Engine(blob, B, seed=1)
eng.reset()
gen = torch.Generator(device='cuda').manual_seed(0)
for t in range(steps):
  eng.step_state(torch.randint(0, eng.num_actions, (B, players), generator=gen, device='cuda', dtype=torch.int32))
full = timeit(eng.render)
eng.set_flags(3 | 16); stores = timeit(eng.render)
eng.set_flags(3 | 32); compute = timeit(eng.render)
extra = {}
for nm, fl in (('neither', 48), ('compose_nocellpass', 32 | 64), ('compose_nofence', 32 | 128), ('neither_nocellpass', 48 | 64), ('compose_allfast', 32 | 256), ('allfast', 256)):
  eng.set_flags((3 if fl > 3 else 0) | fl); extra[nm] = round(timeit(eng.render), 4)
eng.set_flags(3)
grid = eng.grid.view(torch.int16)
occupied = (grid != 0).float().sum(1)  # layers occupied per cell
print(json.dumps({'substrate': name, 'envs': B, 'after_steps': steps, 'render_ms': full, 'stores_only_ms': stores, 'compose_only_ms': compute, 'extra': extra,
                  'mean_layers_per_cell': float(occupied.mean()), 'cells_with_3plus_layers': float((occupied >= 3).float().mean()),
                  'alive': float(eng.avatar_state[:, :, 3].float().mean())}))
