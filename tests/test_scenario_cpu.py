"""Scenario layer host logic against the reference's contract.

Mirrors what `/root/reference/meltingpot/utils/scenarios/scenario_test.py` checks: focal / background
partition of rewards, observations and specs, restriction to permitted observations, action merging,
population resets on episode starts, and the error messages.
"""

import numpy as np
import pytest

from meltingpot_b200 import scenario, shims
from meltingpot_b200 import substrate as substrate_lib

shims.install()
import dm_env  # noqa: E402


class FakeSubstrate:
  """4 players; observation i holds the player index; step echoes the full action as rewards."""

  def __init__(self):
    self.actions = []
    self.closed = False
    self._obs = tuple({'RGB': np.full((2, 2, 3), i, np.uint8), 'SECRET': np.float64(i), 'WORLD.RGB': np.zeros((1,))}
                      for i in range(4))
    self._n = 0

  def reset(self):
    self._n = 0
    return dm_env.TimeStep(dm_env.StepType.FIRST, (0.0,) * 4, 0.0, self._obs)

  def step(self, action):
    self.actions.append(tuple(action))
    self._n += 1
    kind = dm_env.StepType.FIRST if self._n == 3 else dm_env.StepType.MID
    return dm_env.TimeStep(kind, tuple(float(a) for a in action), 1.0, self._obs)

  def observation(self):
    return self._obs

  def action_spec(self):
    return tuple(f'a{i}' for i in range(4))

  def observation_spec(self):
    return tuple({'RGB': f'rgb{i}', 'SECRET': f's{i}', 'WORLD.RGB': 'w'} for i in range(4))

  def reward_spec(self):
    return tuple(f'r{i}' for i in range(4))

  def discount_spec(self):
    return 'd'

  def observables(self):
    return substrate_lib.SubstrateObservables(action=substrate_lib.Subject(), timestep=substrate_lib.Subject(),
                                              events=substrate_lib.Subject())

  def close(self):
    self.closed = True


class FakePopulation:
  def __init__(self):
    self.resets = 0
    self.seen = []
    self.closed = False

  def reset(self):
    self.resets += 1

  def send_timestep(self, timestep):
    self.seen.append(timestep)

  def await_action(self):
    return (70, 71)  # for the two background slots, in slot order

  def close(self):
    self.closed = True


IS_FOCAL = (True, False, True, False)


def _make():
  sub, pop = FakeSubstrate(), FakePopulation()
  return scenario.Scenario(sub, pop, IS_FOCAL, permitted_observations={'RGB', 'WORLD.RGB'}), sub, pop


def test_is_focal_length_is_checked():
  with pytest.raises(ValueError, match='is_focal is length 3 but substrate is 4-player.'):
    scenario.Scenario(FakeSubstrate(), FakePopulation(), (True, False, True), {'RGB'})


def test_specs_are_partitioned_and_restricted():
  sc, _, _ = _make()
  assert sc.action_spec() == ('a0', 'a2')
  assert sc.reward_spec() == ('r0', 'r2')
  assert sc.discount_spec() == 'd'
  assert sc.observation_spec() == ({'RGB': 'rgb0', 'WORLD.RGB': 'w'}, {'RGB': 'rgb2', 'WORLD.RGB': 'w'})


def test_reset_and_step_split_the_timestep():
  sc, sub, pop = _make()
  emitted = []
  sc.observables().timestep.subscribe(emitted.append)
  ts = sc.reset()
  assert pop.resets == 1
  assert len(ts.observation) == 2 and set(ts.observation[0]) == {'RGB', 'WORLD.RGB'}
  assert int(ts.observation[1]['RGB'][0, 0, 0]) == 2  # focal slots are substrate players 0 and 2
  assert set(pop.seen[-1].observation[0]) == {'RGB', 'SECRET', 'WORLD.RGB'}  # bots see everything
  assert float(pop.seen[-1].observation[1]['SECRET']) == 3.0
  ts = sc.step((5, 6))
  assert sub.actions[-1] == (5, 70, 6, 71)  # merged back into substrate order
  assert ts.reward == (5.0, 6.0) and pop.seen[-1].reward == (70.0, 71.0)
  assert emitted[-1] is ts
  sc.step((1, 1))
  assert pop.resets == 1
  ts = sc.step((1, 1))  # third step: the fake substrate starts a new episode
  assert ts.step_type.first() and pop.resets == 2
  assert sc.events() == ()
  assert [int(o['RGB'][0, 0, 0]) for o in sc.observation()] == [0, 2]


def test_wrong_number_of_focal_actions():
  sc, _, _ = _make()
  sc.reset()
  with pytest.raises(ValueError, match='Expected 2 focal actions, got 3.'):
    sc.step((1, 2, 3))


def test_close_closes_both():
  sc, sub, pop = _make()
  done = []
  sc.observables().action.subscribe(on_completed=lambda: done.append(1))
  sc.close()
  assert sub.closed and pop.closed and done == [1]
