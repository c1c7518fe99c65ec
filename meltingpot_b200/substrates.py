"""Locates / builds compiled substrate blobs."""

from __future__ import annotations

import os
from typing import Optional, Sequence

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')

# substrate -> player counts with a committed blob (default roles repeated).
PRECOMPILED = {
    'clean_up': (7,),
    'commons_harvest__open': (7, 16),
    'territory__rooms': (9,),
    'territory__open': (9,),
    'territory__inside_out': (5,),
    'commons_harvest__closed': (7,),
    'commons_harvest__partnership': (7,),
    'coins': (2,),
    'coop_mining': (6,),
}

# Substrates whose CONFIG BUILDER draws from Python's `random` (coins.py:45-84,488: map size, coin colours): the
# reference makes one draw per `build()` call; a compiled blob fixes one draw, made with this seed (policy A.20).
# ('choice' prefabs -- territory__inside_out -- are NOT fixed at compile time: the engine draws them per env and per
# episode, as the reference's prefab_utils.lua:63-65 does at every env build.)
BUILD_SEEDS = {'coins': 0}


def blob_path(name: str, num_players: int) -> str:
  return os.path.join(_DATA, f'{name}__{num_players}p.mpb')


def load_blob(name: str, roles: Optional[Sequence[str]] = None) -> bytes:
  """Returns the compiled blob for `name` with `roles`.

  Uses the committed blob when the roles are the substrate's default role
  repeated; otherwise compiles from a reference checkout (compiler.reference_root()).
  """
  from meltingpot_b200 import compiler  # pylint: disable=g-import-not-at-top
  num_players = len(roles) if roles is not None else None
  if roles is None or len(set(roles)) == 1:
    counts = PRECOMPILED.get(name, ())
    n = num_players if num_players is not None else (counts[0] if counts else None)
    if n is not None and os.path.exists(blob_path(name, n)):
      if roles is None or _is_default_role(name, roles[0]):
        with open(blob_path(name, n), 'rb') as f:
          return f.read()
  if compiler.reference_root() is None:
    raise FileNotFoundError(
        f'no precompiled blob for {name!r} with roles {roles!r} and no Melting Pot '
        'reference checkout to compile from (set MELTINGPOT_REFERENCE_ROOT)')
  return compiler.compile_substrate(name, roles, build_seed=BUILD_SEEDS.get(name))


def _is_default_role(name: str, role: str) -> bool:
  del name
  return role == 'default'
