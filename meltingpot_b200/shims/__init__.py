"""Dependency shims: `dm_env`, `ml_collections`, `immutabledict`, `tree`, `chex`, `reactivex`, `dmlab2d`.

`install()` registers a stand-in module in `sys.modules` for each of these
third-party packages that cannot be imported. The real package always wins.
"""

from __future__ import annotations

import importlib
import sys


def _missing(name: str) -> bool:
  if name in sys.modules:
    return False
  try:
    importlib.import_module(name)
    return False
  except ImportError:
    return True


def install() -> None:
  """Makes `dm_env`, `ml_collections` and `immutabledict` importable."""
  if _missing('dm_env'):
    from meltingpot_b200.shims import dm_env_shim  # pylint: disable=g-import-not-at-top
    dm_env, specs = dm_env_shim.build_modules()
    sys.modules['dm_env'] = dm_env
    sys.modules['dm_env.specs'] = specs
  need_ml = _missing('ml_collections')
  need_imm = _missing('immutabledict')
  if need_ml or need_imm:
    from meltingpot_b200.shims import config_dict_shim  # pylint: disable=g-import-not-at-top
    ml_collections, config_dict, imm = config_dict_shim.build_modules()
    if need_ml:
      sys.modules['ml_collections'] = ml_collections
      sys.modules['ml_collections.config_dict'] = config_dict
    if need_imm:
      sys.modules['immutabledict'] = imm
  from meltingpot_b200.shims import misc_shims  # pylint: disable=g-import-not-at-top
  modules = None
  for name in ('tree', 'chex', 'reactivex'):
    if _missing(name):
      modules = modules or misc_shims.build_modules()
      sys.modules[name] = modules[name]
      if name == 'reactivex':
        sys.modules['reactivex.subject'] = modules['reactivex.subject']
  if _missing('dmlab2d'):
    # The FFI boundary itself: `dmlab2d.Lab2d` / `dmlab2d.Environment` (builder.py:179-187) backed by libmpengine.so.
    from meltingpot_b200 import lab2d_env  # pylint: disable=g-import-not-at-top
    for name, module in lab2d_env.build_modules().items():
      sys.modules[name] = module
