"""Regenerates the committed derived-data blobs under meltingpot_b200/data/.

Needs a reference checkout (default /root/reference, or MELTINGPOT_REFERENCE_ROOT).
The blobs hold numeric tables and RGBA sprite pixels derived from the reference's
substrate configs; no reference source is copied.

  python tools/compile_substrates.py [name[:num_players] ...]
"""
import os
os.environ.setdefault('MELTINGPOT_REFERENCE_ROOT', '/root/reference')  # this tool runs where the checkout is
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meltingpot_b200 import compiler  # noqa: E402
from meltingpot_b200 import substrates  # noqa: E402


def main(argv):
  targets = argv or [f'{n}:{p}' for n, ps in substrates.PRECOMPILED.items() for p in ps]
  for target in targets:
    name, _, players = target.partition(':')
    config = compiler.load_reference_config(name)
    roles = tuple(config.default_player_roles)
    if players:
      roles = (roles[0],) * int(players)
    blob = compiler.compile_substrate(name, roles, build_seed=substrates.BUILD_SEEDS.get(name))
    path = substrates.blob_path(name, len(roles))
    with open(path, 'wb') as f:
      f.write(blob)
    print(f'{path}: {len(blob)} bytes')


if __name__ == '__main__':
  main(sys.argv[1:])
