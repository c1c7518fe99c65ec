"""Compiles variant substrates used only by the tests (needs the reference checkout).

clean_up_clean_river: clean_up with every initially dirty river cell ('F') made clean ('H'), so that
apples grow from the first frame (growth probability 0.05) and random walkers eat them -- this
exercises AppleGrow / Edible / double-entry edge cases that a random policy on the stock map,
whose river only gets dirtier, never reaches.
"""
import os
os.environ.setdefault('MELTINGPOT_REFERENCE_ROOT', '/root/reference')  # this tool runs where the checkout is
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meltingpot_b200 import compiler  # noqa: E402


def main():
  config = compiler.load_reference_config('clean_up')
  roles = tuple(config.default_player_roles)
  settings = compiler._plain(config.lab2d_settings_builder(roles=roles, config=config))
  settings['simulation']['map'] = settings['simulation']['map'].replace('F', 'H')
  blob = compiler.compile_settings(settings, config)
  path = os.path.join(ROOT, 'tests', 'golden', 'clean_up_clean_river__7p.mpb')
  with open(path, 'wb') as f:
    f.write(blob)
  print(path, len(blob))


if __name__ == '__main__':
  main()
