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


def test_exchange_world_of_one_gathers_own_rows(clean_up_blob):
  # The publish path of the state-transition kernel (peer stores + flag + flow control) with the only peer being
  # this rank itself: `gathered` must track timestep_packed on every step, on alternating slots, incl. masked resets.
  import torch
  from meltingpot_b200 import engine
  B = 96
  eng = engine.Engine(clean_up_blob, B, device=0, seed=5)
  ptr, nbytes = eng.exchange_create(0, 1)
  assert nbytes == 256 + 2 * B * (eng.num_players + 2) * 8
  eng.exchange_connect([ptr])
  gen = torch.Generator(device='cuda').manual_seed(0)
  eng.reset()
  slots = set()
  for t in range(40):
    eng.exchange_wait()
    torch.cuda.synchronize()
    slot, step = eng.exchange_slot()
    assert step == t + 1
    slots.add(slot)
    assert torch.equal(eng.gathered_timestep(), eng.timestep_packed), t
    if t == 20:
      mask = torch.zeros(B, dtype=torch.uint8, device='cuda'); mask[::3] = 1
      eng.reset(mask)   # envs outside the mask republish their unchanged rows into the new slot
    else:
      eng.step(torch.randint(0, 9, (B, 7), generator=gen, device='cuda', dtype=torch.int32))
  assert slots == {0, 1}


def test_exchange_two_gpus_one_process(territory_blob):
  # Two ranks in ONE process (peer access enabled directly instead of CUDA IPC): after every step both ranks hold the
  # same stacked rows, equal to the timestep of the unsharded batch.
  import torch
  from meltingpot_b200 import engine
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  B, P = 48, 9
  full = engine.Engine(territory_blob, 2 * B, device=0, seed=7)
  ranks = [engine.Engine(territory_blob, B, device=r, seed=7, env_index_base=r * B) for r in range(2)]
  ptrs = [e.exchange_create(r, 2)[0] for r, e in enumerate(ranks)]
  engine.enable_peer_access(0, 1); engine.enable_peer_access(1, 0)
  for e in ranks:
    e.exchange_connect(ptrs)
  full.reset()
  for e in ranks:
    e.reset()
  gen = torch.Generator().manual_seed(1)
  for t in range(50):
    a = torch.randint(0, 9, (2 * B, P), generator=gen, dtype=torch.int32)
    full.step(a.cuda(0))
    for r, e in enumerate(ranks):
      e.step(a[r * B:(r + 1) * B].contiguous().cuda(r))
    for e in ranks:
      e.exchange_wait()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    want = full.timestep_packed.cpu()
    for e in ranks:
      assert torch.equal(e.gathered_timestep().cpu(), want), t


def test_gather_obs_world_of_one(commons_blob):
  # The renderer's extra (peer) stores with the only peer being this rank: the stacked buffer tracks rgb / world_rgb.
  import torch
  from meltingpot_b200 import engine
  B = 200   # balanced rounds + cooperative tail both exercised (148 CTAs x 4 teams > 200: tail only; see next test for rounds)
  for B in (200, 1300):
    eng = engine.Engine(commons_blob, B, device=0, seed=9)
    ptr, _ = eng.gather_obs_create(0, 1)
    eng.gather_obs_connect([ptr])
    gen = torch.Generator(device='cuda').manual_seed(0)
    eng.reset()
    for t in range(6):
      eng.gather_obs_wait()
      torch.cuda.synchronize()
      rgb, world = eng.gathered_observations()
      assert torch.equal(rgb, eng.rgb) and torch.equal(world, eng.world_rgb), (B, t)
      eng.step(torch.randint(0, eng.num_actions, (B, eng.num_players), generator=gen, device='cuda', dtype=torch.int32))
    eng.gather_obs_enable(False)
    before = eng.gathered_observations()[0].clone()
    eng.step(torch.randint(0, eng.num_actions, (B, eng.num_players), generator=gen, device='cuda', dtype=torch.int32))
    torch.cuda.synchronize()
    assert torch.equal(eng.gathered_observations()[0], before)  # switched off: nothing delivered
    eng.close()


def test_gather_obs_two_gpus_one_process(clean_up_blob):
  import torch
  from meltingpot_b200 import engine
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  B = 300
  ranks = [engine.Engine(clean_up_blob, B, device=r, seed=3, env_index_base=r * B) for r in range(2)]
  ptrs = [e.gather_obs_create(r, 2)[0] for r, e in enumerate(ranks)]
  engine.enable_peer_access(0, 1); engine.enable_peer_access(1, 0)
  for e in ranks:
    e.gather_obs_connect(ptrs)
  for e in ranks:
    e.reset()
  gen = torch.Generator().manual_seed(1)
  for t in range(8):
    for e in ranks:
      e.gather_obs_wait()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    want_rgb = torch.cat([e.rgb.cpu() for e in ranks]); want_world = torch.cat([e.world_rgb.cpu() for e in ranks])
    for e in ranks:
      rgb, world = e.gathered_observations()
      assert torch.equal(rgb.cpu(), want_rgb) and torch.equal(world.cpu(), want_world), t
    a = torch.randint(0, 9, (2 * B, 7), generator=gen, dtype=torch.int32)
    for r, e in enumerate(ranks):
      e.step(a[r * B:(r + 1) * B].contiguous().cuda(r))


def test_sharded_substrate_world_of_one_over_torch_distributed(clean_up_blob):
  # distributed.ShardedSubstrate end to end in a one-rank process group: IPC export of the exchange blocks, connect,
  # stacked timestep and stacked observations equal the local ones.
  import os
  import torch
  import torch.distributed as dist
  from meltingpot_b200 import distributed
  created = False
  if not dist.is_initialized():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
    dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
    created = True
  try:
    sh = distributed.ShardedSubstrate('clean_up', ('default',) * 7, global_num_envs=64, seed=4, device=0)
    assert (sh.env_index_base, sh.local_num_envs) == (0, 64)
    sh.connect(observations=True)
    ts = sh.reset()
    gen = torch.Generator(device='cuda').manual_seed(0)
    for _ in range(5):
      reward, discount, step_type = sh.stacked_timestep()
      rgb, world = sh.stacked_observations()
      torch.cuda.synchronize()
      assert torch.equal(reward, ts.reward) and torch.equal(discount, ts.discount) and torch.equal(step_type, ts.step_type)
      assert torch.equal(rgb, ts.observation['RGB']) and torch.equal(world, ts.observation['WORLD.RGB'])
      ts = sh.step(torch.randint(0, 9, (64, 7), generator=gen, device='cuda', dtype=torch.int32))
    sh.close()
  finally:
    if created:
      dist.destroy_process_group()
