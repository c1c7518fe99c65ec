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
      assert nev[b] <= evs.shape[1], f'events {where}: {nev[b]} events exceed max_events {evs.shape[1]} (the bound of mp_create is wrong)'
      got = sorted(tuple(int(v) for v in row) for row in evs[b][:int(nev[b])])
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


def _event_keys(events, counts):
  """Sorted int64 keys of each env's event rows ([B, M, 3] -> [B, M], unused rows = int64 max)."""
  ev = events.astype(np.int64)
  key = (ev[..., 0] << 44) | ((ev[..., 1] & 0x3fffff) << 22) | (ev[..., 2] & 0x3fffff)
  mask = np.arange(ev.shape[1])[None, :] >= counts[:, None]
  key[mask] = np.iinfo(np.int64).max
  key.sort(axis=1)
  return key


def compare_batch(blob, oracle, num_envs, steps, seed, action_seed=0, pixels_every=10, actions_fn=None,
                  env_index_base=0, threads=None):
  """EVERY env of the batch against the oracle: rewards, discount, step type, scalar observations, avatar state, the
  whole sprite grid and the events on every step, every RGB byte of every env every `pixels_every` steps."""
  import os
  import torch
  from meltingpot_b200 import engine
  threads = threads or os.cpu_count() or 1
  eng = engine.Engine(blob, num_envs, device=0, seed=seed, env_index_base=env_index_base)
  P, A = eng.num_players, eng.num_actions
  bf = eng.buffers
  shapes = dict(P=P, L=int(bf.grid_layers), cells=int(bf.grid_cells), n_scalar=eng.num_scalar_obs,
                rgb=(int(bf.rgb_h), int(bf.rgb_w)), world=(int(bf.world_h), int(bf.world_w)))
  max_ev = int(bf.max_events)
  batch = oracle.OracleBatch(blob, num_envs, seed=seed + env_index_base)  # oracle_batch_create starts episode 0
  rng = np.random.default_rng(action_seed)
  eng.reset()
  stats = dict(rewards=0.0, lasts=0, events=0, pixel_checks=0, envs=num_envs)

  def check(t):
    torch.cuda.synchronize()
    px = (t % pixels_every) == 0
    want = batch.dump(threads, shapes, pixels=px, max_events=max_ev)
    where = f'step {t}'

    def same(name, got, exp):
      if not np.array_equal(got, exp):
        bad = np.argwhere(got != exp)
        raise AssertionError(f'{name} {where}: {len(bad)} mismatches, first at {bad[0].tolist()} (env first): gpu {got[tuple(bad[0])]} oracle {exp[tuple(bad[0])]}')

    same('step_type', eng.step_type.cpu().numpy(), want['step_type'])
    same('discount', eng.discount.cpu().numpy(), want['discount'])
    same('reward', eng.reward.cpu().numpy(), want['reward'])
    if shapes['n_scalar']:
      same('scalar_obs', eng.scalar_obs.cpu().numpy()[:shapes['n_scalar']], want['scalar_obs'][:shapes['n_scalar']])
    same('avatars', eng.avatar_state.cpu().numpy(), want['avatars'])
    same('grid', eng.grid.cpu().numpy().view(np.uint16)[:, :, :shapes['cells']], want['grid'])
    nev = eng.event_count.cpu().numpy()
    same('event count', nev, want['n_events'])
    assert int(nev.max(initial=0)) <= max_ev, f'{where}: {int(nev.max())} events exceed max_events {max_ev}'
    same('events', _event_keys(eng.events.cpu().numpy(), nev), _event_keys(want['events'], want['n_events']))
    if px:
      same('RGB', eng.rgb.cpu().numpy(), want['rgb'])
      same('WORLD.RGB', eng.world_rgb.cpu().numpy(), want['world'])
      stats['pixel_checks'] += 1
    stats['rewards'] += float(want['reward'].sum())
    stats['lasts'] += int((want['step_type'] == 2).sum())
    stats['events'] += int(want['n_events'].sum())

  check(0)
  for t in range(1, steps + 1):
    acts = actions_fn(t, num_envs, P, A, rng) if actions_fn is not None else rng.integers(0, A, size=(num_envs, P))
    acts = np.ascontiguousarray(acts, np.int32)
    eng.step(torch.from_numpy(acts).cuda())
    batch.step_actions(acts, threads)
    check(t)
  eng.close()
  batch.close()
  return stats
