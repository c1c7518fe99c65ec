"""Measures write-only HBM ceilings on the GPU box (for the render-kernel roofline discussion)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meltingpot_b200 import engine, substrates

def timeit(fn, n=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / n

nbytes = 4096 * 283584
x = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
y = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
t = timeit(lambda: x.fill_(7))
print(json.dumps({'test': 'torch fill_ (write only)', 'bytes': nbytes, 'ms': t, 'GBps': nbytes / t / 1e6}))
t = timeit(lambda: y.copy_(x))
print(json.dumps({'test': 'torch copy_ (read+write)', 'bytes': 2 * nbytes, 'ms': t, 'GBps': 2 * nbytes / t / 1e6}))
blob = substrates.load_blob('clean_up')
for flags, name in ((3, 'k_render full'), (3 | 16, 'k_render stores only (no compose)')):
  eng = engine.Engine(blob, 4096, seed=1, flags=flags)
  eng.reset()
  t = timeit(lambda: eng.render())
  _, rb = eng.algorithmic_bytes()
  print(json.dumps({'test': name, 'bytes': rb * 4096, 'ms': t, 'GBps': rb * 4096 / t / 1e6}))
  eng.close()
