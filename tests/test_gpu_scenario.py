"""BatchedScenario on the GPU: the focal / background split is index plumbing around the same engine."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batched_scenario_equals_direct_stepping(clean_up_blob):
  import torch
  from meltingpot_b200 import scenario, substrate
  B, P = 32, 7
  is_focal = (True, True, False, True, False, False, True)
  focal = [i for i, f in enumerate(is_focal) if f]
  background = [i for i, f in enumerate(is_focal) if not f]
  gen = torch.Generator(device='cuda').manual_seed(5)
  seen = []

  def bots(ts):  # background policy: sees unrestricted observations of the 3 background slots
    assert ts.observation['RGB'].shape[:2] == (B, 3) and 'READY_TO_SHOOT' in ts.observation
    a = torch.randint(0, 9, (B, 3), generator=gen, device='cuda', dtype=torch.int32)
    seen.append(a.clone())
    return a

  sc = scenario.BatchedScenario(substrate.BatchedSubstrate(clean_up_blob, B, seed=9), bots, is_focal,
                                permitted_observations={'RGB', 'COLLECTIVE_REWARD'})
  direct = substrate.BatchedSubstrate(clean_up_blob, B, seed=9)
  ts = sc.reset()
  ref = direct.reset()
  assert set(ts.observation) == {'RGB', 'COLLECTIVE_REWARD'} and ts.observation['RGB'].shape[:2] == (B, 4)
  rng = np.random.default_rng(0)
  for t in range(120):
    fa = torch.from_numpy(np.ascontiguousarray(rng.integers(0, 9, (B, 4)), np.int32)).cuda()
    ts = sc.step(fa)
    full = torch.zeros((B, P), dtype=torch.int32, device='cuda')
    full[:, focal] = fa
    full[:, background] = seen[-1]
    ref = direct.step(full)
    assert torch.equal(ts.reward, ref.reward[:, focal])
    assert torch.equal(sc.background_timestep.reward, ref.reward[:, background])
    assert torch.equal(ts.step_type, ref.step_type)
    if t % 10 == 0:
      assert torch.equal(ts.observation['RGB'], ref.observation['RGB'][:, focal])
      assert torch.equal(ts.observation['COLLECTIVE_REWARD'], ref.observation['COLLECTIVE_REWARD'])
  with pytest.raises(ValueError, match='Expected 4 focal actions'):
    sc.step(torch.zeros((B, 5), dtype=torch.int32, device='cuda'))
  with pytest.raises(ValueError, match='is_focal is length 3'):
    scenario.BatchedScenario(direct, bots, (True, False, True), {'RGB'})
