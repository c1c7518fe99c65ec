"""Shared GPU-vs-oracle comparison loop (used by the -m gpu tests)."""

import numpy as np


def compare_rollout(blob, oracle, num_envs, steps, seed, check_envs=None, action_seed=0,
                    pixels_every=1, actions_fn=None, env_index_base=0):
  import torch
  from meltingpot_b200 import engine
  eng = engine.Engine(blob, num_envs, device=0, seed=seed, env_index_base=env_index_base)
  P, A = eng.num_players, eng.num_actions
  check_envs = list(range(num_envs)) if check_envs is None else list(check_envs)
  envs = {b: oracle.OracleEnv(blob, seed + env_index_base + b) for b in check_envs}
  rng = np.random.default_rng(action_seed)
  eng.reset()
  for e in envs.values():
    e.reset()
  stats = dict(rewards=0.0, lasts=0, zaps=0, cleaned=0, eaten=0, events=0)
  code_of = {v: k for k, v in oracle.EVENT_NAMES.items()}

  def check(t, acts):
    torch.cuda.synchronize()
    rew = eng.reward.cpu().numpy(); disc = eng.discount.cpu().numpy(); st = eng.step_type.cpu().numpy()
    sc = eng.scalar_obs.cpu().numpy(); av = eng.avatar_state.cpu().numpy()
    grid = eng.grid.cpu().numpy().view(np.uint16)
    nev = eng.event_count.cpu().numpy(); evs = eng.events.cpu().numpy()
    do_px = (t % pixels_every) == 0
    if do_px:
      idx = torch.as_tensor(check_envs, device='cuda')
      rgb = eng.rgb[idx].cpu().numpy(); world = eng.world_rgb[idx].cpu().numpy()
    for i, b in enumerate(check_envs):
      e = envs[b]
      where = f'step {t} env {b}'
      assert e.step_type() == st[b], f'step_type {where}: {e.step_type()} vs {st[b]}'
      assert e.discount() == disc[b], f'discount {where}'
      np.testing.assert_array_equal(e.rewards(), rew[b], err_msg=f'reward {where}')
      np.testing.assert_array_equal(e.scalar_obs().T, sc[:e.n_scalar, b, :], err_msg=f'scalar obs {where}')
      np.testing.assert_array_equal(e.avatars(), av[b], err_msg=f'avatars {where}')
      og = e.grid()
      gg = grid[b][:, :og.shape[1]]
      if not np.array_equal(og, gg):
        bad = np.argwhere(og != gg)
        raise AssertionError(f'grid {where}: first diffs (layer, cell) {bad[:8].tolist()} '
                             f'oracle {og[tuple(bad[0])]} gpu {gg[tuple(bad[0])]}')
      if do_px:
        np.testing.assert_array_equal(e.rgb(), rgb[i], err_msg=f'RGB {where}')
        np.testing.assert_array_equal(e.world_rgb(), world[i], err_msg=f'WORLD.RGB {where}')
      stats['rewards'] += float(rew[b].sum())
      stats['lasts'] += int(st[b] == 2)
      want = sorted((code_of[name], a, b2) for name, a, b2 in e.events())
      assert nev[b] == len(want), f'event count {where}: oracle {len(want)} gpu {nev[b]}'
      got = sorted(tuple(int(v) for v in row) for row in evs[b][:min(int(nev[b]), evs.shape[1])])
      if len(want) <= evs.shape[1]:  # (beyond max_events the engine keeps an unspecified subset)
        assert got == want, f'events {where}: oracle {want} gpu {got}'
      stats['events'] += len(want)
      for name, _, _ in e.events():
        key = {'zap': 'zaps', 'player_cleaned': 'cleaned', 'edible_consumed': 'eaten'}.get(name)
        if key:
          stats[key] += 1

  check(-1, None)
  for t in range(steps):
    if actions_fn is not None:
      acts = actions_fn(t, num_envs, P, A, rng)
    else:
      acts = rng.integers(0, A, size=(num_envs, P))
    acts = np.ascontiguousarray(acts, np.int32)
    eng.step(torch.from_numpy(acts).cuda())
    for b, e in envs.items():
      e.step(acts[b])
    check(t, acts)
  eng.close()
  return stats
