"""The reference's own substrate test (`meltingpot/testing/substrates.py:22-68`, used by
`meltingpot/substrate_test.py:24-47` for every substrate) restated against this package's dm_env surface."""

import numpy as np
import pytest

from meltingpot_b200 import substrates

pytestmark = pytest.mark.gpu

CASES = sorted(substrates.PRECOMPILED)


@pytest.mark.parametrize('name', CASES)
def test_step_matches_specs(name):
  from meltingpot_b200 import substrate
  config = substrate.get_config(name)
  with substrate.build_from_config(config, roles=config.default_player_roles) as env:
    env.reset()
    action = [spec.maximum for spec in env.action_spec()]
    timestep = env.step(action)
    env.discount_spec().validate(timestep.discount)
    reward_spec = env.reward_spec()
    assert len(reward_spec) == len(timestep.reward)
    for r, spec in zip(timestep.reward, reward_spec):
      spec.validate(r)
    observation_specs = env.observation_spec()
    assert len(observation_specs) == len(timestep.observation) == len(config.default_player_roles)
    for observation, spec in zip(timestep.observation, observation_specs):
      assert set(spec) == set(observation)
      for key in spec:
        spec[key].validate(observation[key])
    # what the config promises (substrate_test.py: the specs come from the config)
    for key, spec in config.timestep_spec.observation.items():
      assert observation_specs[0][key] == spec
    assert env.action_spec()[0] == config.action_spec


@pytest.mark.parametrize('name', CASES)
def test_seed_determinism_and_episode_variation(name):
  # builder_test.py:47-106: same seed -> same first frame; another seed or the next episode -> another frame.
  from meltingpot_b200 import substrate
  config = substrate.get_config(name)
  roles = config.default_player_roles
  def first(seed, resets=1):
    with substrate.build_from_config(config, roles=roles, env_seed=seed) as env:
      for _ in range(resets):
        ts = env.reset()
      return ts.observation[0]['WORLD.RGB'].copy()
  a = first(123)
  np.testing.assert_array_equal(a, first(123))
  assert not np.array_equal(a, first(124))
  assert not np.array_equal(a, first(123, resets=2))
