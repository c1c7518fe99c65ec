"""bench.py contract checks that need no GPU: the reference arm's JSON line, and the product arm refusing to run."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_exactly_one_json_line_with_the_contract_keys():
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-seconds', '1'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-500:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  line = json.loads(lines[0])
  assert line['impl'] == 'reference' and line['metric'] == 'env_steps_per_sec' and line['unit'] == 'env-steps/s'
  assert line['higher_is_better'] is True and line['scaling'] == 'weak' and line['vs_baseline'] is None
  assert line['value'] > 0 and line['steps'] == 1 and line['warmup'] == 0 and line['dtype'] == 'u8'
  assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
  assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
  assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
  assert 'workload' in line['config']


def test_reference_arm_is_silent_on_other_ranks():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--ref-seconds', '1'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
  assert out.returncode == 0 and out.stdout.strip() == ''


def test_product_arm_refuses_to_run_without_a_gpu():
  import torch
  if torch.cuda.is_available():
    return
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert out.returncode != 0 and out.stdout.strip() == ''
  assert 'no CPU path' in out.stderr or 'no CUDA device' in out.stderr
