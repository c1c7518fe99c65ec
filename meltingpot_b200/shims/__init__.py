"""Dependency shims: `dm_env`, `ml_collections`, `immutabledict`.

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
