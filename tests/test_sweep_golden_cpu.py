"""Committed blobs against what the reference's Python configs say (fixture made by tools/make_sweep_golden.py).

Runs without a reference checkout: the fixture holds the reference-side facts, the blobs are in-tree.
"""

import hashlib
import json
import os

import numpy as np
import pytest

from meltingpot_b200 import blob as mpb
from meltingpot_b200 import substrate, substrates

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, 'tests', 'golden', 'sweep_reference_golden.json')) as _f:
  GOLDEN = json.load(_f)['substrates']


def test_every_committed_blob_has_a_golden_entry():
  want = {f'{n}:{p}' for n, ps in substrates.PRECOMPILED.items() for p in ps}
  assert want == set(GOLDEN)


@pytest.mark.parametrize('key', sorted(GOLDEN))
def test_blob_matches_the_reference_config(key):
  name, players = key.split(':')
  players = int(players)
  g = GOLDEN[key]
  blob = substrates.load_blob(name, ('default',) * players)
  assert hashlib.sha256(blob).hexdigest() == g['blob_sha256']  # the blob the fixture was generated next to
  sec = mpb.unpack(blob)
  info = json.loads(mpb.section_text(sec, 'info_json'))
  meta = sec['meta']
  assert [int(meta[1]), int(meta[2])] == g['map_size']
  assert int(meta[4]) == g['num_players'] == info['num_players']
  assert int(meta[5]) == g['sprite_size'] and int(meta[7]) == g['max_episode_length_frames']
  assert int(meta[6]) == {'BOUNDED': 0, 'TORUS': 1}[g['topology']]
  assert info['action_set'] == g['action_set'] and int(meta[19]) == len(g['action_set'])
  assert info['individual_observation_names'] == g['individual_observation_names']
  assert info['global_observation_names'] == g['global_observation_names']
  assert info['rgb_shape'] == g['observation_specs']['RGB']['shape']
  assert info['world_rgb_shape'] == g['observation_specs']['WORLD.RGB']['shape']
  assert info['valid_roles'] == g['valid_roles'] and info['default_player_roles'] == g['default_player_roles']
  # one object per map character that has a prefab (compound prefabs add more), plus scene and avatars
  mapped = sum(n for ch, n in g['map_census'].items() if ch in g['char_prefab_map_keys'])
  assert int(meta[8]) >= mapped + players


@pytest.mark.parametrize('key', sorted(GOLDEN))
def test_public_config_follows_the_golden(key):
  name, players = key.split(':')
  g = GOLDEN[key]
  if int(players) != len(g['default_player_roles']):
    pytest.skip('non-default player count')
  cfg = substrate.get_config(name)
  assert list(cfg.default_player_roles) == g['default_player_roles']
  assert sorted(cfg.valid_roles) == g['valid_roles']
  assert cfg.action_spec.num_values == len(g['action_set'])
  for obs, spec in g['observation_specs'].items():
    assert list(cfg.timestep_spec.observation[obs].shape) == spec['shape']
    assert str(np.dtype(cfg.timestep_spec.observation[obs].dtype)) == spec['dtype']
