"""Generates tests/golden/sweep_reference_golden.json by IMPORTING the Python reference configs.

For every substrate with a committed blob: what the reference's own Python data says about it (API metadata,
map census, lab2d settings summary, component census of the object list) next to the SHA-256 of the committed
blob. tests/test_sweep_golden_cpu.py checks the committed blobs against it without a reference checkout (the
GPU box has none). Substrates with build-time randomness use substrates.BUILD_SEEDS.

  python tools/make_sweep_golden.py
"""
import os
os.environ.setdefault('MELTINGPOT_REFERENCE_ROOT', '/root/reference')  # this tool runs where the checkout is
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meltingpot_b200 import compiler, substrates  # noqa: E402


def describe(name, players):
  config = compiler.load_reference_config(name)
  roles = (tuple(config.default_player_roles)[0],) * players
  seed = substrates.BUILD_SEEDS.get(name)
  state = random.getstate()
  try:
    if seed is not None:
      random.seed(seed)
    settings = compiler._plain(config.lab2d_settings_builder(roles=roles, config=config))  # pylint: disable=protected-access
  finally:
    random.setstate(state)
  sim = settings['simulation']
  rows = [r for r in sim['map'].strip('\n').split('\n')]
  census = {}
  for row in rows:
    for ch in row:
      census[ch] = census.get(ch, 0) + 1
  comps = {}
  for prefab in sim['prefabs'].values():
    for c in prefab['components']:
      comps[c['component']] = comps.get(c['component'], 0) + 1
  avatar_comps = sorted({c['component'] for go in sim.get('gameObjects', []) for c in go['components']})
  spec = config.timestep_spec
  return {
      'num_players': players,
      'level_name': settings['levelName'],
      'topology': settings.get('topology', 'BOUNDED'),
      'max_episode_length_frames': int(settings['maxEpisodeLengthFrames']),
      'sprite_size': int(settings['spriteSize']),
      'map_size': [max(len(r) for r in rows), len(rows)],
      'map_census': census,
      'char_prefab_map_keys': sorted(str(k) for k in sim['charPrefabMap']),
      'prefab_component_census': comps,
      'avatar_components': avatar_comps,
      'action_set': [dict(a) for a in config.action_set],
      'individual_observation_names': list(config.individual_observation_names),
      'global_observation_names': list(config.global_observation_names),
      'observation_specs': {k: {'shape': list(v.shape), 'dtype': str(v.dtype)} for k, v in spec.observation.items()},
      'valid_roles': sorted(config.valid_roles),
      'default_player_roles': list(config.default_player_roles),
      'build_seed': seed,
      'blob_sha256': hashlib.sha256(substrates.load_blob(name, ('default',) * players)).hexdigest(),
  }


def main():
  golden = {'generated_by': 'tools/make_sweep_golden.py (imports the reference Python configs under /root/reference)',
            'substrates': {}}
  for name, counts in sorted(substrates.PRECOMPILED.items()):
    for players in counts:
      golden['substrates'][f'{name}:{players}'] = describe(name, players)
  path = os.path.join(ROOT, 'tests', 'golden', 'sweep_reference_golden.json')
  with open(path, 'w') as f:
    json.dump(golden, f, indent=1, sort_keys=True)
  print(path, len(golden['substrates']), 'entries')


if __name__ == '__main__':
  main()
