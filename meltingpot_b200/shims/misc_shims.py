"""Stand-ins for `tree`, `chex`, `reactivex` (used only when the real packages are absent).

The reference's builder and wrapper stack import these for three small things: `tree.map_structure`
(`/root/reference/meltingpot/utils/substrates/builder.py:51`), `chex.dataclass`
(`utils/substrates/substrate.py:33`, `wrappers/observables.py:32`) and `reactivex.subject.Subject`
(`wrappers/observables_wrapper.py:38-40`). Each is restated here from its public contract.
"""

from __future__ import annotations

import dataclasses
import types
from typing import Any, List


# ---- tree ------------------------------------------------------------------------------------
def map_structure(func, *structures):
  """dm-tree's map_structure for the containers configs hold: dicts (key order kept), lists, tuples, leaves."""
  first = structures[0]
  if isinstance(first, dict):
    return type(first)((k, map_structure(func, *(s[k] for s in structures))) for k in first)
  if isinstance(first, tuple) and hasattr(first, '_fields'):
    return type(first)(*(map_structure(func, *xs) for xs in zip(*structures)))
  if isinstance(first, (list, tuple)):
    return type(first)(map_structure(func, *xs) for xs in zip(*structures))
  return func(*structures)


def flatten(structure) -> List[Any]:
  out: List[Any] = []
  if isinstance(structure, dict):
    for k in sorted(structure):
      out += flatten(structure[k])
  elif isinstance(structure, (list, tuple)):
    for v in structure:
      out += flatten(v)
  else:
    out.append(structure)
  return out


# ---- chex ------------------------------------------------------------------------------------
def chex_dataclass(cls=None, *, frozen=False, **unused):
  def wrap(c):
    return dataclasses.dataclass(frozen=frozen)(c)
  return wrap if cls is None else wrap(cls)


# ---- reactivex ---------------------------------------------------------------------------------
class Observable:
  """Minimal hot observable: subscribe(on_next, on_error, on_completed) or subscribe(observer)."""

  def __init__(self):
    self._observers: List[Any] = []

  def subscribe(self, on_next=None, on_error=None, on_completed=None, **unused):
    if on_next is not None and not callable(on_next):  # observer object
      observer = on_next
      entry = (getattr(observer, 'on_next', None), getattr(observer, 'on_error', None),
               getattr(observer, 'on_completed', None))
    else:
      entry = (on_next, on_error, on_completed)
    self._observers.append(entry)
    return _Disposable(self, entry)

  def __class_getitem__(cls, item):  # annotations like reactivex.Observable[dm_env.TimeStep]
    return cls


class _Disposable:
  def __init__(self, source, entry):
    self._source, self._entry = source, entry

  def dispose(self):
    if self._entry in self._source._observers:  # pylint: disable=protected-access
      self._source._observers.remove(self._entry)  # pylint: disable=protected-access


class Subject(Observable):
  def on_next(self, value):
    for fn, _, _ in list(self._observers):
      if fn:
        fn(value)

  def on_error(self, error):
    for _, fn, _ in list(self._observers):
      if fn:
        fn(error)

  def on_completed(self):
    for _, _, fn in list(self._observers):
      if fn:
        fn()
    self._observers.clear()


def build_modules():
  """Returns {module name: module} for tree, chex, reactivex, reactivex.subject."""
  tree = types.ModuleType('tree')
  tree.map_structure = map_structure
  tree.flatten = flatten
  chex = types.ModuleType('chex')
  chex.dataclass = chex_dataclass
  reactivex = types.ModuleType('reactivex')
  reactivex.Observable = Observable
  subject = types.ModuleType('reactivex.subject')
  subject.Subject = Subject
  reactivex.subject = subject
  reactivex.Subject = Subject
  return {'tree': tree, 'chex': chex, 'reactivex': reactivex, 'reactivex.subject': subject}
