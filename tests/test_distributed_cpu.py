"""world_size-2 gloo test of the env-sharding + scalar all-gather logic (no GPU needed)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meltingpot_b200 import distributed


def test_shard_envs_partitions_the_global_range():
  spans = [distributed.shard_envs(4096, r, 8) for r in range(8)]
  assert spans[0] == (0, 512) and spans[7] == (3584, 512)
  assert sum(c for _, c in spans) == 4096 and all(b == i * 512 for i, (b, _) in enumerate(spans))
  with pytest.raises(ValueError):
    distributed.shard_envs(10, 0, 3)
  with pytest.raises(ValueError):
    distributed.shard_envs(8, 2, 2)


def _worker(rank, world, port, ret):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    base, count = distributed.shard_envs(8, rank, world)
    p = 7
    env_ids = torch.arange(base, base + count, dtype=torch.float64)
    reward = env_ids[:, None] * 10 + torch.arange(p, dtype=torch.float64)[None, :]
    discount = torch.where(env_ids % 2 == 0, 1.0, 0.0).to(torch.float64)
    step_type = (env_ids.to(torch.int64) % 3)
    r, d, s = distributed.gather_timestep_scalars(reward, discount, step_type)
    ok = (r.shape == (8, p) and torch.equal(r[:, 0], torch.arange(8, dtype=torch.float64) * 10)
          and torch.equal(d, torch.tensor([1., 0.] * 4, dtype=torch.float64))
          and torch.equal(s, torch.arange(8) % 3) and s.dtype == torch.int64)
    rgb = torch.full((count, 2, 3), float(rank), dtype=torch.float32)
    stacked = distributed.all_gather_stacked(rgb)
    ok = ok and stacked.shape == (8, 2, 3) and torch.equal(stacked[:, 0, 0], torch.tensor([0.] * 4 + [1.] * 4))
    ret[rank] = bool(ok)
  finally:
    dist.destroy_process_group()


def test_gloo_world_size_two_gathers_in_global_env_order():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  manager = mp.Manager()
  ret = manager.dict()
  mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
  assert ret[0] and ret[1]
