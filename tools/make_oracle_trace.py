"""Writes tests/golden/clean_up_oracle_trace.json: a regression trace of the CPU oracle.

SELF-GENERATED (not reference-derived): it freezes the oracle's current behaviour under the engine
policy ledger so that accidental semantic drift in oracle or kernels is caught. Real-DMLab2D traces
can only be captured where dmlab2d is installed (tools/capture_reference_golden.py).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meltingpot_b200 import substrates  # noqa: E402
from oracle import binding  # noqa: E402

PROBS = [0.05, 0.25, 0.05, 0.05, 0.05, 0.1, 0.1, 0.05, 0.3]


def trace(blob, seed, steps, action_seed, every=25):
  env = binding.OracleEnv(blob, seed)
  env.reset()
  rng = np.random.default_rng(action_seed)
  h = hashlib.sha256()
  out, total = [], np.zeros(env.P)
  for t in range(steps):
    a = rng.choice(9, size=env.P, p=PROBS)
    st = env.step(a)
    total += env.rewards()
    h.update(env.rewards().tobytes()); h.update(env.avatars().tobytes()); h.update(env.scalar_obs().tobytes())
    if (t + 1) % every == 0:
      h.update(env.rgb().tobytes()); h.update(env.world_rgb().tobytes())
      out.append({'step': t + 1, 'step_type': int(st), 'sha256': h.hexdigest(), 'return': total.tolist(),
                  'dirt': env.counters()['dirt']})
  return out


def main():
  blob = substrates.load_blob('clean_up')
  with open(os.path.join(ROOT, 'tests', 'golden', 'clean_up_clean_river__7p.mpb'), 'rb') as f:
    clean = f.read()
  rec = {'substrate': 'clean_up', 'note': 'self-generated oracle regression trace; NOT a DMLab2D golden',
         'seed': 42, 'action_seed': 7, 'action_probs': PROBS, 'checkpoints': trace(blob, 42, 600, 7),
         'clean_river_checkpoints': trace(clean, 43, 400, 8)}
  path = os.path.join(ROOT, 'tests', 'golden', 'clean_up_oracle_trace.json')
  with open(path, 'w') as f:
    json.dump(rec, f, indent=1)
  print(path, rec['checkpoints'][-1], rec['clean_river_checkpoints'][-1])


if __name__ == '__main__':
  main()
