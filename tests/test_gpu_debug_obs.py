"""POSITION / ORIENTATION / LAYER / zap matrix against the oracle."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('fixture', ['clean_up_blob', 'territory_blob'])
def test_debug_observations_equal_the_oracle(fixture, oracle, request):
  import torch
  from meltingpot_b200 import engine
  blob = request.getfixturevalue(fixture)
  B, seed = 6, 17
  eng = engine.Engine(blob, B, seed=seed)
  envs = [oracle.OracleEnv(blob, seed + b) for b in range(B)]
  eng.reset()
  for e in envs:
    e.reset()
  rng = np.random.default_rng(5)
  zaps = 0
  for t in range(120):
    obs = {k: v.cpu().numpy() for k, v in eng.debug_observations().items()}
    for b, e in enumerate(envs):
      av = e.avatars()
      assert np.array_equal(obs['POSITION'][b], av[:, :2]) and np.array_equal(obs['ORIENTATION'][b], av[:, 2])
      assert np.array_equal(obs['LAYER'][b], e.layer_view()), (t, b)
      want = np.zeros((e.P, e.P), np.int32)
      for name, a, c in e.events():
        if name == 'zap':
          want[a - 1, c - 1] += 1
      assert np.array_equal(obs['ZAP_MATRIX'][b], want)
      zaps += int(want.sum())
    acts = rng.integers(0, eng.num_actions, size=(B, eng.num_players)).astype(np.int32)
    eng.step(torch.from_numpy(acts).cuda())
    for b, e in enumerate(envs):
      e.step(acts[b])
  assert zaps > 0
  assert obs['POSITION'].dtype == np.int32 and obs['ORIENTATION'].dtype == np.int32


def test_position_and_orientation_specs_match_the_reference_table():
  from meltingpot_b200 import specs
  assert specs.OBSERVATION['POSITION'].shape == (2,) and specs.OBSERVATION['POSITION'].dtype == np.int32
  assert specs.OBSERVATION['ORIENTATION'].shape == () and specs.OBSERVATION['ORIENTATION'].dtype == np.int32
