"""Cost of the cross-GPU timestep exchange on ONE GPU (world of one: the peer stores land in local memory): step time with
the exchange off / publishing only / publishing + the consumer-side wait kernel on a side stream or on the main stream."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meltingpot_b200 import engine, substrates

blob = substrates.load_blob('clean_up', ('default',) * 7)
B, K, W = 4096, 400, 20
for mode in ('off', 'publish', 'publish+wait_side', 'publish+wait_main'):
  eng = engine.Engine(blob, B, seed=1)
  if mode != 'off':
    ptr, _ = eng.exchange_create(0, 1)
    eng.exchange_connect([ptr])
  gen = torch.Generator(device='cuda').manual_seed(0)
  acts = torch.randint(0, 9, (K + W, B, 7), generator=gen, device='cuda', dtype=torch.int32)
  stream = torch.cuda.current_stream()
  side = torch.cuda.Stream()
  ev = torch.cuda.Event()
  def step(t):
    if mode == 'publish+wait_side':
      eng.step(acts[t])
      ev.record(stream)
      side.wait_event(ev)
      eng.exchange_wait(side)
    elif mode == 'publish+wait_main':
      eng.step(acts[t])
      eng.exchange_wait(stream)
    else:
      eng.step(acts[t])
  eng.reset()
  for t in range(W):
    step(t)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for t in range(W, W + K):
    step(t)
  stream.wait_stream(side)
  e1.record()
  torch.cuda.synchronize()
  print(json.dumps({'mode': mode, 'ms_per_step': e0.elapsed_time(e1) / K}), flush=True)
  eng.close()
