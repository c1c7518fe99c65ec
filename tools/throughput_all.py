"""Device-resident throughput of every supported substrate at the BASELINE.json config sizes.

Writes one JSON line per config (steps/s, render / step kernel split, roofline fraction of the render kernel).
Not the bench line (bench.py measures BASELINE.json's metric on clean_up); kept under profiles/ as context.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meltingpot_b200 import engine, substrates

CONFIGS = [
    ('clean_up', 7, 4096),
    ('commons_harvest__open', 16, 8192),
    ('commons_harvest__open', 7, 4096),
    ('territory__rooms', 9, 2048),
    ('territory__rooms', 9, 4096),
    ('territory__open', 9, 2048),
    ('territory__inside_out', 5, 4096),
    ('commons_harvest__closed', 7, 4096),
    ('commons_harvest__partnership', 7, 4096),
    ('coins', 2, 8192),
    ('coop_mining', 6, 4096),
]
PEAK = 6561.6
if os.path.exists('MEASURED_PEAKS.json'):
  PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']

for name, players, B in CONFIGS:
  blob = substrates.load_blob(name, ('default',) * players)
  eng = engine.Engine(blob, B, seed=1)
  A = eng.num_actions
  gen = torch.Generator(device='cuda').manual_seed(0)
  K, W = 300, 30
  acts = torch.randint(0, A, (K + W, B, players), generator=gen, device='cuda', dtype=torch.int32)
  eng.reset()
  for t in range(W):
    eng.step(acts[t])
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for t in range(W, W + K):
    eng.step(acts[t])
  e1.record(); torch.cuda.synchronize()
  total_ms = e0.elapsed_time(e1) / K
  evs = []
  for t in range(40):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(); eng.step_state(acts[W + t]); b.record(); eng.render(); c.record()
    evs.append((a, b, c))
  torch.cuda.synchronize()
  step_ms = sum(a.elapsed_time(b) for a, b, c in evs) / len(evs)
  render_ms = sum(b.elapsed_time(c) for a, b, c in evs) / len(evs)
  algo, rbytes = eng.algorithmic_bytes()
  print(json.dumps({'substrate': name, 'players': players, 'envs': B, 'ms_per_step': total_ms,
                    'env_steps_per_sec': B / total_ms * 1e3, 'agent_steps_per_sec': B * players / total_ms * 1e3,
                    'step_kernel_ms': step_ms, 'render_kernel_ms': render_ms,
                    'render_GBps': rbytes * B / render_ms / 1e6, 'render_frac_of_peak': rbytes * B / render_ms / 1e6 / PEAK,
                    'render_bytes_per_env': rbytes, 'whole_step_bytes_per_env': algo, 'render_plan': eng.render_plan()}), flush=True)
  if name == 'clean_up':  # secondary line of SURVEY.md section 8d config 2: WORLD.RGB off
    eng.set_flags(engine.MP_FLAG_RENDER_PLAYERS)
    for t in range(W):
      eng.step(acts[t])
    torch.cuda.synchronize()
    e0.record()
    for t in range(W, W + K):
      eng.step(acts[t])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(json.dumps({'substrate': name, 'players': players, 'envs': B, 'world_rgb': False, 'ms_per_step': ms,
                      'env_steps_per_sec': B / ms * 1e3, 'agent_steps_per_sec': B * players / ms * 1e3}), flush=True)
  eng.close()
