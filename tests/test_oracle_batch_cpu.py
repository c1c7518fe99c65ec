"""The whole-batch checker API of the oracle (oracle_batch_step_actions / oracle_batch_dump) against per-env calls."""

import numpy as np
import pytest


@pytest.mark.parametrize('fixture', ['clean_up_blob', 'territory_blob', 'commons_blob'])
def test_batch_dump_equals_per_env_calls(fixture, oracle, request):
  blob = request.getfixturevalue(fixture)
  B, seed, steps = 6, 40, 60
  envs = [oracle.OracleEnv(blob, seed + b) for b in range(B)]
  for e in envs:
    e.reset()
  batch = oracle.OracleBatch(blob, B, seed=seed)
  e0 = envs[0]
  shapes = dict(P=e0.P, L=e0.L, cells=e0.W * e0.H, n_scalar=e0.n_scalar, rgb=e0.rgb_shape[:2], world=e0.world_shape[:2])
  rng = np.random.default_rng(3)
  code_of = {v: k for k, v in oracle.EVENT_NAMES.items()}
  total_events = 0
  for t in range(steps):
    if t:
      acts = rng.integers(0, e0.n_actions, size=(B, e0.P)).astype(np.int32)
      batch.step_actions(acts, 3)
      for b, e in enumerate(envs):
        e.step(acts[b])
    d = batch.dump(2, shapes, pixels=(t % 8 == 0), max_events=128)
    for b, e in enumerate(envs):
      assert np.array_equal(d['reward'][b], e.rewards()) and d['discount'][b] == e.discount() and d['step_type'][b] == e.step_type()
      assert np.array_equal(d['scalar_obs'][:e.n_scalar, b, :], e.scalar_obs().T)
      assert np.array_equal(d['avatars'][b], e.avatars()) and np.array_equal(d['grid'][b], e.grid())
      want = sorted((code_of[n], a, c) for n, a, c in e.events())
      assert d['n_events'][b] == len(want)
      assert [tuple(int(v) for v in r) for r in d['events'][b][:len(want)]] == want
      total_events += len(want)
      if 'rgb' in d:
        assert np.array_equal(d['rgb'][b], e.rgb()) and np.array_equal(d['world'][b], e.world_rgb())
  assert total_events > 0
