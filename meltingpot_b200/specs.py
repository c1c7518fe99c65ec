"""dm_env specs for substrate timesteps (mirrors the helpers substrates' configs use).

Same names and semantics as `/root/reference/meltingpot/utils/substrates/specs.py:26-160`
(`STEP_TYPE`, `DISCOUNT`, `REWARD`, `OBSERVATION`, `rgb`, `float64`, `action`, `timestep`),
built on `dm_env` (or the stand-in from `meltingpot_b200.shims`).
"""

from __future__ import annotations

from typing import Mapping, Optional

import numpy as np

from meltingpot_b200 import shims

shims.install()
import dm_env  # noqa: E402  pylint: disable=g-import-not-at-top,g-bad-import-order

STEP_TYPE = dm_env.specs.BoundedArray(
    shape=(), dtype=np.int64, minimum=min(dm_env.StepType),
    maximum=max(dm_env.StepType), name='step_type')
DISCOUNT = dm_env.specs.BoundedArray(
    shape=(), dtype=np.float64, minimum=0, maximum=1, name='discount')
REWARD = dm_env.specs.Array(shape=(), dtype=np.float64, name='reward')
OBSERVATION = {
    'READY_TO_SHOOT': dm_env.specs.Array(shape=(), dtype=np.float64, name='READY_TO_SHOOT'),
    'RGB': dm_env.specs.Array(shape=(88, 88, 3), dtype=np.uint8, name='RGB'),
    'POSITION': dm_env.specs.Array(shape=(2,), dtype=np.int32, name='POSITION'),  # specs.py:39-44 of the reference
    'ORIENTATION': dm_env.specs.Array(shape=(), dtype=np.int32, name='ORIENTATION'),
}
_ACTION = dm_env.specs.DiscreteArray(num_values=1, dtype=np.int64, name='action')


def float64(*shape: int, name: Optional[str] = None):
  return dm_env.specs.Array(shape=shape, dtype=np.float64, name=name)


def rgb(height: int, width: int, name: Optional[str] = 'RGB'):
  return OBSERVATION['RGB'].replace(shape=(height, width, 3), name=name)


def action(num_actions: int):
  return _ACTION.replace(num_values=num_actions)


def timestep(observation_spec: Mapping[str, 'dm_env.specs.Array']) -> 'dm_env.TimeStep':
  """Spec of a single player's timestep; observation spec names follow their keys."""
  observation = {}
  for name, spec in observation_spec.items():
    observation[name] = spec.replace(name=name)
  return dm_env.TimeStep(step_type=STEP_TYPE, discount=DISCOUNT, reward=REWARD,
                         observation=observation)
