"""Minimal stand-ins for `ml_collections.config_dict` and `immutabledict`.

The reference's substrate configs build an `ml_collections.ConfigDict`
(`/root/reference/meltingpot/configs/substrates/clean_up.py:806-838`) and
`specs.py` uses `immutabledict` (`/root/reference/meltingpot/utils/substrates/
specs.py:36`). Neither package exists in this image; these classes provide the
subset of behaviour the reference's config modules exercise: attribute and item
access, nesting of plain dicts, `lock()/unlock()/unlocked()`, `to_dict()`.
"""

from __future__ import annotations

import contextlib
import types
from collections.abc import Mapping


class ConfigDict:
  """Attribute-access dict with lock support (subset of ml_collections)."""

  def __init__(self, initial_dictionary=None):
    object.__setattr__(self, '_fields', {})
    object.__setattr__(self, '_locked', False)
    if initial_dictionary is not None:
      if isinstance(initial_dictionary, ConfigDict):
        initial_dictionary = initial_dictionary._fields
      for key, value in dict(initial_dictionary).items():
        self[key] = value

  @staticmethod
  def _wrap(value):
    if isinstance(value, dict):
      return ConfigDict(value)
    return value

  # -- locking ---------------------------------------------------------------
  @property
  def is_locked(self):
    return self._locked

  def lock(self):
    object.__setattr__(self, '_locked', True)
    for value in self._fields.values():
      if isinstance(value, ConfigDict):
        value.lock()
    return self

  def unlock(self):
    object.__setattr__(self, '_locked', False)
    for value in self._fields.values():
      if isinstance(value, ConfigDict):
        value.unlock()
    return self

  @contextlib.contextmanager
  def unlocked(self):
    was_locked = self._locked
    if was_locked:
      self.unlock()
    try:
      yield self
    finally:
      if was_locked:
        self.lock()

  # -- mapping protocol ------------------------------------------------------
  def __setitem__(self, key, value):
    if '.' in str(key):
      raise KeyError('ConfigDict does not accept dots in field names: %r' % key)
    if self._locked and key not in self._fields:
      raise KeyError('This ConfigDict is locked, you have to unlock it before '
                     'adding new fields: %r' % key)
    self._fields[key] = self._wrap(value)

  def __getitem__(self, key):
    return self._fields[key]

  def __delitem__(self, key):
    if self._locked:
      raise KeyError('This ConfigDict is locked, cannot delete %r' % key)
    del self._fields[key]

  def __setattr__(self, name, value):
    try:
      self[name] = value
    except KeyError as e:
      raise AttributeError(str(e)) from e

  def __getattr__(self, name):
    try:
      return object.__getattribute__(self, '_fields')[name]
    except KeyError as e:
      raise AttributeError(name) from e

  def __delattr__(self, name):
    try:
      del self[name]
    except KeyError as e:
      raise AttributeError(name) from e

  def __contains__(self, key):
    return key in self._fields

  def __iter__(self):
    return iter(self._fields)

  def __len__(self):
    return len(self._fields)

  def keys(self):
    return self._fields.keys()

  def values(self):
    return self._fields.values()

  def items(self):
    return self._fields.items()

  def get(self, key, default=None):
    return self._fields.get(key, default)

  def update(self, *other, **kwargs):
    for mapping in other:
      for key, value in dict(mapping).items():
        self[key] = value
    for key, value in kwargs.items():
      self[key] = value

  def to_dict(self):
    out = {}
    for key, value in self._fields.items():
      out[key] = value.to_dict() if isinstance(value, ConfigDict) else value
    return out

  def copy_and_resolve_references(self):
    return ConfigDict(self.to_dict())

  def __eq__(self, other):
    if isinstance(other, ConfigDict):
      return self._fields == other._fields
    if isinstance(other, dict):
      return self.to_dict() == other
    return NotImplemented

  def __repr__(self):
    return 'ConfigDict(%r)' % (self.to_dict(),)

  def __deepcopy__(self, memo):
    import copy  # pylint: disable=g-import-not-at-top
    out = ConfigDict()
    for key, value in self._fields.items():
      out._fields[key] = copy.deepcopy(value, memo)
    object.__setattr__(out, '_locked', self._locked)
    return out


Mapping.register(ConfigDict)


class immutabledict(dict):  # pylint: disable=invalid-name
  """A dict that refuses mutation (subset of the `immutabledict` package)."""

  def _immutable(self, *args, **kwargs):
    raise TypeError('immutabledict does not support mutation')

  __setitem__ = _immutable
  __delitem__ = _immutable
  clear = _immutable
  pop = _immutable
  popitem = _immutable
  setdefault = _immutable
  update = _immutable

  def __hash__(self):
    return hash(frozenset(self.items()))

  def __repr__(self):
    return 'immutabledict(%s)' % dict.__repr__(self)


def build_modules():
  """Returns (ml_collections, ml_collections.config_dict, immutabledict)."""
  ml_collections = types.ModuleType('ml_collections')
  config_dict = types.ModuleType('ml_collections.config_dict')
  config_dict.ConfigDict = ConfigDict
  ml_collections.config_dict = config_dict
  ml_collections.ConfigDict = ConfigDict
  ml_collections.__meltingpot_b200_shim__ = True
  imm = types.ModuleType('immutabledict')
  imm.immutabledict = immutabledict
  imm.__meltingpot_b200_shim__ = True
  return ml_collections, config_dict, imm
