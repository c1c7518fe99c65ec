"""GPU (CUDA engine through the C ABI) == CPU oracle, bit for bit, on identical seeds/actions."""

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


def test_clean_up_random_rollout(clean_up_blob, oracle):
  stats = parity.compare_rollout(clean_up_blob, oracle, num_envs=16, steps=300, seed=1)
  assert stats['zaps'] > 0 and stats['cleaned'] > 0


def _cleaning_policy(t, B, P, A, rng):
  # Mostly clean/move so that the river gets clean, apples grow and get eaten.
  probs = np.array([0.05, 0.25, 0.05, 0.05, 0.05, 0.1, 0.1, 0.05, 0.3])
  return rng.choice(A, size=(B, P), p=probs)


def test_clean_up_cleaning_policy_grows_and_eats_apples(clean_up_blob, oracle):
  stats = parity.compare_rollout(clean_up_blob, oracle, num_envs=8, steps=700, seed=77,
                                 actions_fn=_cleaning_policy, pixels_every=7)
  assert stats['cleaned'] > 0


def test_clean_up_sharding_invariance(clean_up_blob, oracle):
  # env b of a shard that starts at env_index_base behaves like global env base+b.
  parity.compare_rollout(clean_up_blob, oracle, num_envs=4, steps=40, seed=5, env_index_base=1000)
