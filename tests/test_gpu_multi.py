"""Two-GPU sharding: each device steps its shard; together they equal the single-device batch."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_shards_equal_one_batch(territory_blob):
  import torch
  from meltingpot_b200 import distributed, engine
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  B, P = 64, 9
  full = engine.Engine(territory_blob, B, device=0, seed=11)
  shards = []
  for r in range(2):
    base, count = distributed.shard_envs(B, r, 2)
    shards.append(engine.Engine(territory_blob, count, device=r, seed=11, env_index_base=base))
  full.reset()
  for s in shards:
    s.reset()
  gen = torch.Generator().manual_seed(0)
  for t in range(60):
    a = torch.randint(0, 9, (B, P), generator=gen, dtype=torch.int32)
    full.step(a.cuda(0))
    for r, s in enumerate(shards):
      s.step(a[r * 32:(r + 1) * 32].contiguous().cuda(r))
  torch.cuda.synchronize(0); torch.cuda.synchronize(1)
  for r, s in enumerate(shards):
    sl = slice(r * 32, (r + 1) * 32)
    assert np.array_equal(full.rgb[sl].cpu().numpy(), s.rgb.cpu().numpy())
    assert np.array_equal(full.world_rgb[sl].cpu().numpy(), s.world_rgb.cpu().numpy())
    assert np.array_equal(full.reward[sl].cpu().numpy(), s.reward.cpu().numpy())
    assert np.array_equal(full.timestep_packed[sl].cpu().numpy(), s.timestep_packed.cpu().numpy())
