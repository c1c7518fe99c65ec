"""Generates tests/golden/*_reference_golden.json by IMPORTING the Python reference.

Run in a container that has the reference checkout (default /root/reference). The fixture pins
what the substrate compiler must reproduce from the reference's own Python data: action set,
observation names and specs, the ASCII-map census, palette arithmetic (shapes.get_palette /
scale_color) and the sprite pixels the reference's shapes + palettes produce.
"""
import os
os.environ.setdefault('MELTINGPOT_REFERENCE_ROOT', '/root/reference')  # this tool runs where the checkout is
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meltingpot_b200 import compiler  # noqa: E402


def sprite_rows(text, palette):
  rows = text.strip('\n').split('\n')
  return [[list(int(c) for c in (tuple(palette[ch]) + (255,))[:4]) for ch in row] for row in rows]


def main(name='clean_up'):
  config = compiler.load_reference_config(name)
  import importlib
  module = importlib.import_module(f'meltingpot.configs.substrates.{name}')
  shapes = importlib.import_module('meltingpot.utils.substrates.shapes')
  colors = importlib.import_module('meltingpot.utils.substrates.colors')
  roles = tuple(config.default_player_roles)
  settings = config.lab2d_settings_builder(roles=roles, config=config)
  rows = [r for r in settings['simulation']['map'].strip('\n').split('\n')]
  census = {}
  for row in rows:
    for ch in row:
      census[ch] = census.get(ch, 0) + 1
  spec = config.timestep_spec
  golden = {
      'substrate': name,
      'generated_by': 'tools/make_golden_from_reference.py (imports /root/reference Python configs)',
      'num_players': len(roles),
      'action_set': [dict(a) for a in config.action_set],
      'individual_observation_names': list(config.individual_observation_names),
      'global_observation_names': list(config.global_observation_names),
      'observation_specs': {k: {'shape': list(v.shape), 'dtype': str(v.dtype)} for k, v in spec.observation.items()},
      'reward_dtype': str(spec.reward.dtype), 'discount_dtype': str(spec.discount.dtype),
      'step_type_dtype': str(spec.step_type.dtype),
      'action_num_values': int(config.action_spec.num_values), 'action_dtype': str(config.action_spec.dtype),
      'valid_roles': sorted(config.valid_roles), 'default_player_roles': list(config.default_player_roles),
      'map_size': [len(rows[0]), len(rows)], 'map_census': census,
      'max_episode_length_frames': int(settings['maxEpisodeLengthFrames']),
      'sprite_size': int(settings['spriteSize']), 'topology': settings['topology'],
      'self_palette': {k: list(v) for k, v in shapes.get_palette(colors.human_readable[0]).items()},
      'avatar2_palette': {k: list(v) for k, v in shapes.get_palette(colors.human_readable[1]).items()},
      'scale_color_samples': [[list(c), f, list(shapes.scale_color(c, f, 255))]
                              for c in [(45, 110, 220), (245, 130, 0), (255, 255, 255)] for f in (0.55, 0.75, 1.25)],
      'cute_avatar_n_rows_self': sprite_rows(shapes.CUTE_AVATAR_N, shapes.get_palette(colors.human_readable[0])),
      'wall_rows': sprite_rows(shapes.WALL, {'*': (95, 95, 95, 255), '&': (100, 100, 100, 255),
                                              '@': (109, 109, 109, 255), '#': (152, 152, 152, 255)}),
  }
  blob = compiler.compile_substrate(name, roles)
  golden['blob_sha256'] = hashlib.sha256(blob).hexdigest()
  path = os.path.join(ROOT, 'tests', 'golden', f'{name}_reference_golden.json')
  with open(path, 'w') as f:
    json.dump(golden, f, indent=1, sort_keys=True)
  print(path)


if __name__ == '__main__':
  main(*sys.argv[1:])
