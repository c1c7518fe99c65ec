"""Captures golden traces (and the true step rate) from the REAL reference: dm-meltingpot + dmlab2d.

CANNOT RUN IN THIS REPO'S SANDBOX (dmlab2d is not installable there); it is shipped so that parity
against real DMLab2D can be audited on a machine that has `pip install dm-meltingpot`:

  python tools/capture_reference_golden.py --substrate clean_up --seed 1 --steps 300 --out golden.json
  python tools/capture_reference_golden.py --substrate clean_up --time 30

The trace holds, per step: actions, rewards, discount, step type, sha256 of every RGB observation and
the first frame in full. RNG streams differ by design (DESIGN.md policy A.16: mt19937_64 vs Philox),
so traces are compared distributionally / on RNG-free prefixes (e.g. the first frame's static layers,
movement and beam geometry under scripted actions), not bit-for-bit.
"""
import argparse
import hashlib
import json
import time

import numpy as np


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--substrate', default='clean_up')
  ap.add_argument('--seed', type=int, default=1)
  ap.add_argument('--steps', type=int, default=300)
  ap.add_argument('--out', default='reference_golden.json')
  ap.add_argument('--time', type=float, default=0.0, help='seconds to time uniform-random stepping instead')
  args = ap.parse_args()
  from meltingpot import substrate  # the real reference  # pylint: disable=g-import-not-at-top
  from meltingpot.utils.substrates import builder  # pylint: disable=g-import-not-at-top
  config = substrate.get_config(args.substrate)
  roles = config.default_player_roles
  settings = config.lab2d_settings_builder(roles=roles, config=config)
  rng = np.random.default_rng(1234)
  n_actions = config.action_spec.num_values
  if args.time > 0:
    with substrate.build(args.substrate, roles=roles) as env:
      env.reset()
      n, t0 = 0, time.perf_counter()
      while time.perf_counter() - t0 < args.time:
        ts = env.step(rng.integers(0, n_actions, len(roles)))
        n += 1
        if ts.last():
          env.reset()
      dt = time.perf_counter() - t0
    print(json.dumps({'substrate': args.substrate, 'env_steps_per_sec_one_process': n / dt, 'steps': n}))
    return
  env = substrate.build_from_config(config, roles=roles) if not hasattr(builder, 'builder') else substrate.build(
      args.substrate, roles=roles)
  del settings
  ts = env.reset()
  trace = {'substrate': args.substrate, 'seed': args.seed, 'first_world_rgb': ts.observation[0]['WORLD.RGB'].tolist(),
           'steps': []}
  for _ in range(args.steps):
    actions = rng.integers(0, n_actions, len(roles)).tolist()
    ts = env.step(actions)
    h = hashlib.sha256()
    for obs in ts.observation:
      h.update(np.ascontiguousarray(obs['RGB']).tobytes())
    h.update(np.ascontiguousarray(ts.observation[0]['WORLD.RGB']).tobytes())
    trace['steps'].append({'actions': actions, 'reward': [float(r) for r in ts.reward], 'discount': float(ts.discount),
                           'step_type': int(ts.step_type), 'rgb_sha256': h.hexdigest()})
  env.close()
  with open(args.out, 'w') as f:
    json.dump(trace, f)
  print(args.out)


if __name__ == '__main__':
  main()
