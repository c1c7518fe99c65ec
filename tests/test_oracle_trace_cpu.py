"""Regression: the oracle still reproduces its committed (self-generated) trace."""

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def test_oracle_reproduces_committed_trace(clean_up_blob, oracle):
  import make_oracle_trace
  with open(os.path.join(ROOT, 'tests', 'golden', 'clean_up_oracle_trace.json')) as f:
    rec = json.load(f)
  got = make_oracle_trace.trace(clean_up_blob, rec['seed'], rec['checkpoints'][-1]['step'], rec['action_seed'])
  assert got == rec['checkpoints']
  with open(os.path.join(ROOT, 'tests', 'golden', 'clean_up_clean_river__7p.mpb'), 'rb') as f:
    clean = f.read()
  got = make_oracle_trace.trace(clean, 43, rec['clean_river_checkpoints'][-1]['step'], 8)
  assert got == rec['clean_river_checkpoints']
  assert sum(rec['clean_river_checkpoints'][-1]['return']) > 0  # this one exercises growth and eating
