#!/usr/bin/env python
"""Benchmark: env-steps/s of the batched clean_up hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--envs B]

A "step" is one pass of the hot path (state transition + all observations rendered) over one
batch of `--envs` clean_up instances per GPU under uniform-random actions.
  value  device-resident throughput: actions already in HBM, outputs left in HBM.
  e2e    the same metric through the host-buffer C-ABI call mp_step_host (pinned host actions in,
         every observation copied back to pinned host memory, each step).
  roofline  the render kernel's achieved HBM bandwidth (algorithmic bytes / CUDA-event time).
  cpu_baseline  the C oracle (a port of the reference semantics; DMLab2D itself cannot run here).
Under torchrun (N > 1) every rank steps its own shard of envs; scalars are all-gathered per step.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _ROOT)

SUBSTRATE = 'clean_up'
METRIC = 'env_steps_per_sec'
UNIT = 'env-steps/s'


def _peaks():
  path = os.path.join(_ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


def _traffic_per_launch(num_envs):
  """dram bytes per render launch from the committed ncu capture, scaled to this batch."""
  path = os.path.join(_ROOT, 'profiles', 'render_traffic.json')
  if not os.path.exists(path):
    return None
  with open(path) as f:
    rec = json.load(f)
  return rec['dram_bytes_per_env'] * num_envs


class ClockSampler(threading.Thread):
  """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

  QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index = index
    self.samples = []
    self._halt = threading.Event()

  def run(self):
    while not self._halt.is_set():
      try:
        out = subprocess.run(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                              '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
          self.samples.append([s.strip() for s in out.split(',')])
      except Exception:  # pylint: disable=broad-except
        pass
      self._halt.wait(0.05)

  def stop(self):
    self._halt.set()
    self.join(timeout=5)
    if not self.samples:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
    mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = [n for i, n in enumerate(names) if any(s[2 + i] == 'Active' for s in self.samples if len(s) > 2 + i)]
    return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': reasons, 'samples': len(self.samples)}


def cpu_baseline(blob, seconds=12.0):
  """Times the oracle on all host cores over a bounded sample; returns the cpu_baseline object."""
  from oracle import binding as oracle_binding
  oracle_binding.build()
  cores = os.cpu_count() or 1
  batch = oracle_binding.OracleBatch(blob, cores, seed=1)
  t0 = time.perf_counter()
  n = batch.step_random(100, cores)
  rate = n / (time.perf_counter() - t0)
  steps = max(100, int(rate * seconds / cores))
  t0 = time.perf_counter()
  n = batch.step_random(steps, cores)
  dt = time.perf_counter() - t0
  batch.close()
  return {'value': n / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
          'sample': f'{cores} clean_up envs (one per thread) x {steps} steps, uniform-random actions, '
                    f'all 7 RGB + WORLD.RGB rendered each step, {dt:.1f} s'}


def run_reference(args, rank, world):
  """--impl reference: the CPU implementation of the path (oracle port) on all host cores."""
  if rank != 0:
    return
  from meltingpot_b200 import substrates
  from oracle import binding as oracle_binding
  oracle_binding.build()
  blob = substrates.load_blob(SUBSTRATE)
  cores = os.cpu_count() or 1
  envs = cores
  per_step = 16  # one bench "step" = 16 env-steps on each of `cores` envs (bounded sample)
  batch = oracle_binding.OracleBatch(blob, envs, seed=1)
  for _ in range(args.warmup):
    batch.step_random(per_step, cores)
  t0 = time.perf_counter()
  total = 0
  for _ in range(args.steps):
    total += batch.step_random(per_step, cores)
  dt = time.perf_counter() - t0
  batch.close()
  value = total / dt
  sample = (f'{envs} envs x {per_step} env-steps per bench step, {args.steps} steps, one persistent env per thread, '
            f'full rendering')
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
      'config': {'workload': f'{SUBSTRATE}, 7 players, uniform-random actions, CPU oracle port of the reference path '
                             '(DMLab2D itself is not installable here)', 'envs': envs},
      'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
      'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'agent_steps_per_sec': value * 7,
  }
  _emit(line)


def run_b200(args, rank, world, local_rank):
  import torch
  from meltingpot_b200 import engine, substrates
  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no CUDA device; the B200 engine has no CPU path')
  torch.cuda.set_device(local_rank)
  dist = None
  if world > 1:
    import torch.distributed as dist  # pylint: disable=g-import-not-at-top
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  blob = substrates.load_blob(SUBSTRATE)
  B = args.envs
  eng = engine.Engine(blob, B, device=local_rank, seed=1, env_index_base=rank * B)
  P, A = eng.num_players, eng.num_actions
  algo_bytes, render_bytes = eng.algorithmic_bytes()
  dev = torch.device('cuda', local_rank)
  gen = torch.Generator(device=dev).manual_seed(1234 + rank)
  K, Wm = args.steps, args.warmup
  actions = torch.randint(0, A, (K + Wm, B, P), generator=gen, device=dev, dtype=torch.int32)
  stream = torch.cuda.current_stream(dev)
  gathered = None
  side = None
  if world > 1:
    gathered = torch.empty((world * B, P + 2), dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev)

  def one_step(t):
    if world == 1:
      eng.step(actions[t])
      return
    # The scalar timestep (reward, discount, step type: one packed f64 buffer written by the step
    # kernel) is all-gathered on a side stream while the render kernel runs on the main stream.
    stream.wait_stream(side)            # last step's gather has finished reading timestep_packed
    eng.step_state(actions[t])
    side.wait_stream(stream)
    with torch.cuda.stream(side):
      dist.all_gather_into_tensor(gathered, eng.timestep_packed)
    eng.render()

  eng.reset()
  for t in range(Wm):
    one_step(t)
  torch.cuda.synchronize()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  launches0 = eng.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record(stream)
  for t in range(Wm, Wm + K):
    one_step(t)
  if side is not None:
    stream.wait_stream(side)
  ev1.record(stream)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  elapsed_ms = ev0.elapsed_time(ev1)
  launches = eng.launch_count() - launches0
  clocks = sampler.stop() if rank == 0 else None
  if world > 1:
    tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax.item())
  value = world * B * K / (elapsed_ms * 1e-3)

  # ---- render kernel alone (roofline), CUDA events around each launch ----------------------
  n_r = min(K, 50)
  evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_r)]
  for i in range(n_r):
    eng.step_state(actions[Wm + i])
    evs[i][0].record(stream)
    eng.render()
    evs[i][1].record(stream)
  torch.cuda.synchronize()
  render_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
  peak, peak_src = _peaks()
  achieved = render_bytes * B / (render_ms * 1e-3) / 1e9

  # ---- end to end through the host-buffer C-ABI call ----------------------------------------
  n_e = max(3, min(K, args.e2e_steps))
  host_out = eng.make_host_outputs()
  host_actions = actions[:n_e + 2].cpu().pin_memory()
  d2h = sum(t.numel() * t.element_size() for t in host_out.values())
  h2d = host_actions[0].numel() * host_actions[0].element_size()
  eng.step_host(host_actions[0], host_out)
  eng.step_host(host_actions[1], host_out)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  t0 = time.perf_counter()
  for i in range(n_e):
    eng.step_host(host_actions[2 + i], host_out)  # synchronises the stream before returning
  e2e_s = time.perf_counter() - t0
  if world > 1:
    tmax = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    e2e_s = float(tmax.item())
  e2e_value = world * B * n_e / e2e_s
  # The same call with NULL image pointers: host actions in, rewards / discounts / step types / scalar observations
  # out, images left in HBM for a GPU-resident consumer. Reported next to `e2e`, not instead of it.
  scalars_out = {k: v for k, v in host_out.items() if k not in ('rgb', 'world_rgb')}
  d2h_scalars = sum(t.numel() * t.element_size() for t in scalars_out.values())
  n_s = max(n_e, min(K, 200))
  host_actions_s = actions[:n_s].cpu().pin_memory()
  eng.step_host(host_actions_s[0], scalars_out)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  t0 = time.perf_counter()
  for i in range(n_s):
    eng.step_host(host_actions_s[i], scalars_out)
  e2e_scalars_s = time.perf_counter() - t0
  if world > 1:
    tmax = torch.tensor([e2e_scalars_s], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    e2e_scalars_s = float(tmax.item())
  e2e_scalars_value = world * B * n_s / e2e_scalars_s

  if rank == 0:
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': elapsed_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8', 'data': 'synthetic',
        'config': {
            'workload': f'{SUBSTRATE}, 7 players, {B} batched envs per GPU, uniform-random actions, '
                        'WORLD.RGB on (BASELINE.json configs[1])',
            'envs_per_gpu': B, 'players': P, 'global_envs': world * B,
            'cache': 'per-step working set 1.2 GB of freshly written observations >> 126 MB L2; no explicit flush',
            'multi_gpu': 'env shards, no data-path collective; per-step NCCL all-gather of the packed reward/discount/step_type buffer on a side stream, overlapped with rendering' if world > 1 else 'single GPU',
        },
        'agent_steps_per_sec': value * P,
        'gpu_launches': launches,
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'steps': n_e, 'api': 'mp_step_host (C ABI, pinned host buffers, all observations copied back)'},
        'e2e_scalars_only': {'value': e2e_scalars_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h_scalars,
                             'steps': n_s, 'api': 'mp_step_host with NULL image pointers (images stay in HBM)'},
        'roofline': {'bound': 'hbm', 'kernel': 'k_render', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': _traffic_per_launch(B), 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': render_bytes * B, 'ms_per_launch': render_ms,
                     'whole_step_algorithmic_bytes_per_env': algo_bytes,
                     'whole_step_frac': algo_bytes * B * K / (elapsed_ms * 1e-3) / 1e9 / peak},
    }
    if not args.no_cpu_baseline and world == 1:
      line['cpu_baseline'] = cpu_baseline(blob)
    _emit(line)
  eng.close()
  if world > 1:
    dist.destroy_process_group()


_STDOUT_FD = []  # the real stdout, saved by main() while fd 1 points at stderr


def _emit(line):
  sys.stdout.flush()
  if _STDOUT_FD:
    os.dup2(_STDOUT_FD[0], 1)
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=2000)
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--envs', type=int, default=4096, help='env instances per GPU')
  ap.add_argument('--e2e-steps', type=int, default=30)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  # Exactly one JSON line on stdout: native libraries (NCCL prints its version banner there) write to the stdout
  # file descriptor, so everything but the final line is sent to stderr.
  sys.stdout.flush()
  _STDOUT_FD.append(os.dup(1))
  os.dup2(2, 1)
  if args.impl == 'reference':
    run_reference(args, rank, world)
  else:
    run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
  main()
