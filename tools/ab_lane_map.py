"""A/B of the renderer's lane -> cell dealing (default whole-cell order / scattered colouring / plain) and the two ceilings
(stores without compositing, compositing without stores)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meltingpot_b200 import engine, substrates

CONFIGS = [('clean_up', 7, 4096), ('commons_harvest__open', 16, 8192), ('territory__rooms', 9, 2048), ('coins', 2, 2048)]
for name, players, B in CONFIGS:
  blob = substrates.load_blob(name, ('default',) * players)
  row = {'substrate': name, 'envs': B}
  for label, extra in (('cells', 0), ('scatter', 1 << 10), ('plain', 1 << 9), ('stores_only', 16), ('compose_only', 32)):
    eng = engine.Engine(blob, B, seed=1, flags=3 | extra)
    gen = torch.Generator(device='cuda').manual_seed(0)
    K, W = 200, 20
    acts = torch.randint(0, eng.num_actions, (K + W, B, players), generator=gen, device='cuda', dtype=torch.int32)
    eng.reset()
    for t in range(W):
      eng.step(acts[t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(W, W + K):
      eng.step(acts[t])
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / K
    evs = []
    for t in range(50):
      eng.step_state(acts[W + t])
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(); eng.render(); b.record()
      evs.append((a, b))
    torch.cuda.synchronize()
    row[label] = {'ms_per_step': step_ms, 'render_ms': sum(a.elapsed_time(b) for a, b in evs) / len(evs)}
    eng.close()
  print(json.dumps(row), flush=True)
