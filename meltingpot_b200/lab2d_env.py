"""`dmlab2d.Lab2d` / `dmlab2d.Environment` over libmpengine.so: the reference's FFI hop, served by the B200 engine.

The reference reaches its hot path through exactly two calls
(`/root/reference/meltingpot/utils/substrates/builder.py:179-187`):

    env_raw = dmlab2d.Lab2d(_DMLAB2D_ROOT, lab2d_settings_dict)      # flattened "a.b.1.c" -> str settings
    dmlab2d.Environment(env=env_raw, observation_names=env_raw.observation_names(), seed=seed)

and drives the result through `Lab2dWrapper` (`wrappers/base.py:26-84`): flat observation dicts keyed
"{i}.RGB" / "{i}.REWARD" / "WORLD.RGB", action dicts keyed "{i}.move" ..., dm_env TimeSteps. This module provides
those two classes with that behaviour, so the reference's UNMODIFIED builder.py, wrapper stack and Substrate class
run on top of the engine when it is registered as `dmlab2d` (`meltingpot_b200.shims.install()` does so when the
real dmlab2d is absent). The flattened settings are un-flattened and compiled to an MPB blob (`compiler.py`); the
state transition and the rendering run in the CUDA kernels through the host-buffer C-ABI calls; there is no CPU path.

`Environment` talks to a small backend object (reset / step / outputs / events). The product backend is
`EngineBackend` (ctypes -> libmpengine.so). Tests may install another factory in `BACKEND_FACTORY` -- that is how
the reference wrapper stack is exercised on the CPU oracle in this repo's CPU test-suite; nothing in this package
does.
"""

from __future__ import annotations

import abc
import itertools
import re
import types
from typing import Any, Callable, Dict, List, Mapping, Optional, Sequence

import numpy as np

from meltingpot_b200 import blob as blob_lib

_INT = re.compile(r'^[+-]?\d+$')
_FLOAT = re.compile(r'^[+-]?(\d+\.\d*|\.\d+|\d+)([eE][+-]?\d+)?$|^[+-]?(inf|nan)$')
_NEVER_LISTS = ('charPrefabMap',)  # dicts whose keys may legitimately be the digits "1", "2", ...


# ---------------------------------------------------------------------------------------------------------------------
# Settings: dmlab2d.settings_helper.flatten_args and its inverse
# ---------------------------------------------------------------------------------------------------------------------
def flatten_args(args_in: Mapping[str, Any]) -> Dict[str, Any]:
  """Nested dict / list settings -> flat {"a.b.1.c": value} with 1-based list indices (dmlab2d.settings_helper)."""
  out: Dict[str, Any] = {}

  def walk(prefix, value):
    if hasattr(value, 'to_dict') and not isinstance(value, dict):
      value = value.to_dict()
    if isinstance(value, Mapping):
      for k, v in value.items():
        walk(f'{prefix}{k}.', v)
    elif isinstance(value, (list, tuple)):
      for i, v in enumerate(value):
        walk(f'{prefix}{i + 1}.', v)
    else:
      out[prefix[:-1]] = value

  for key, value in args_in.items():
    walk(f'{key}.', value)
  return out


def _parse_leaf(text: Any) -> Any:
  if not isinstance(text, str):
    return text
  if text == 'True':
    return True
  if text == 'False':
    return False
  if text == 'None':
    return None
  if _INT.match(text):
    return int(text)
  if _FLOAT.match(text):
    return float(text)
  return text


def unflatten_args(flat: Mapping[str, Any]) -> Dict[str, Any]:
  """Inverse of `flatten_args` followed by `str()` (builder.py:55-67): flat string settings -> nested dicts / lists."""
  root: Dict[str, Any] = {}
  for key, value in flat.items():
    parts = key.split('.')
    node = root
    for part in parts[:-1]:
      node = node.setdefault(part, {})
    node[parts[-1]] = _parse_leaf(value)

  def listify(node, name=''):
    if not isinstance(node, dict):
      return node
    node = {k: listify(v, k) for k, v in node.items()}
    keys = list(node)
    if keys and name not in _NEVER_LISTS and all(_INT.match(k) for k in keys):
      idx = sorted(int(k) for k in keys)
      if idx == list(range(1, len(idx) + 1)):
        return [node[str(i)] for i in idx]
    return node

  return listify(root)


# ---------------------------------------------------------------------------------------------------------------------
# The raw action interface: dmlab2d exposes every primitive action field of every avatar
# ---------------------------------------------------------------------------------------------------------------------
class ActionSpace:
  """The avatars' primitive action fields (`Avatar:addActions`, avatar_library.lua:205-216) and the full table of
  their combinations. The engine looks discrete ids up in a table inside the blob; at this (pre-DiscreteActionWrapper)
  boundary every combination the spec allows must be steppable, so the blob is compiled with the full product and
  an action dict is turned into its index by mixed radix."""

  def __init__(self, settings: Mapping[str, Any]):
    from meltingpot_b200 import compiler  # pylint: disable=g-import-not-at-top
    avatar = None
    for go in settings['simulation'].get('gameObjects', []):
      for comp in go['components']:
        if comp['component'] == 'Avatar':
          avatar = comp['kwargs']
          break
      if avatar is not None:
        break
    if avatar is None:
      raise ValueError('settings hold no Avatar game object (builder.maybe_build_and_add_avatar_objects adds them)')
    self.order: List[str] = list(avatar.get('actionOrder', ['move', 'turn']))
    spec = avatar['actionSpec']
    self.bounds = [(int(spec[name]['min']), int(spec[name]['max']), int(spec[name].get('default', 0))) for name in self.order]
    for name in self.order:
      if name not in compiler.ACTION_FIELDS:
        raise NotImplementedError(f'action field {name!r}')
    self.table: List[Dict[str, int]] = []
    for combo in itertools.product(*[range(lo, hi + 1) for lo, hi, _ in self.bounds]):
      self.table.append(dict(zip(self.order, combo)))

  def index(self, values: Sequence[int]) -> int:
    """Row of `table` holding these field values (given in `order`)."""
    idx = 0
    for v, (lo, hi, _) in zip(values, self.bounds):
      v = int(v)
      if not lo <= v <= hi:
        raise ValueError(f'action value {v} outside [{lo}, {hi}]')
      idx = idx * (hi - lo + 1) + (v - lo)
    return idx


class _PseudoConfig:
  """What `compiler.compile_settings` reads from a substrate config, derived from the settings alone."""

  def __init__(self, settings: Mapping[str, Any], actions: ActionSpace):
    from meltingpot_b200 import compiler  # pylint: disable=g-import-not-at-top
    self.action_set = tuple(actions.table)
    names = ['RGB']
    seen = set()
    for go in settings['simulation'].get('gameObjects', []):
      for comp in go['components']:
        kw = comp.get('kwargs', {}) or {}
        if comp['component'] == 'ReadyToShootObservation':
          seen.add('READY_TO_SHOOT')
        elif comp['component'] == 'AvatarMetricReporter':
          for metric in kw.get('metrics', []):
            seen.add(metric['name'])
    names += [n for n in compiler.SCALAR_OBS if n in seen]
    self.individual_observation_names = names
    self.global_observation_names = ['WORLD.RGB']
    self.valid_roles = frozenset()
    self.default_player_roles = ()


# ---------------------------------------------------------------------------------------------------------------------
# dmlab2d.Lab2d
# ---------------------------------------------------------------------------------------------------------------------
class Lab2d:
  """Stands in for `dmlab2d.Lab2d(runfiles_path, settings)`: holds the level description, compiled for the engine."""

  def __init__(self, runfiles_path: str, settings: Mapping[str, str]):
    from meltingpot_b200 import compiler  # pylint: disable=g-import-not-at-top
    del runfiles_path  # the Lua game scripts are not used: the level's semantics live in the CUDA kernels
    self.flat_settings = dict(settings)
    nested = unflatten_args(self.flat_settings)
    self.env_seed = int(nested.pop('env_seed', 0) or 0)
    level = str(nested.get('levelName', ''))
    nested['levelName'] = level.rsplit('/', 1)[-1]  # builder.locate_and_overwrite_level_directory prefixed the directory
    self.settings = nested
    self.actions = ActionSpace(nested)
    # ('choice' prefabs are left to the engine: drawn per env and episode from the env's RNG key, as the reference
    # draws them with the env's random stream at every env build)
    self.blob = compiler.compile_settings(nested, _PseudoConfig(nested, self.actions))
    self.info = blob_lib.unpack(self.blob)
    meta = self.info['meta']
    self.num_players = int(meta[4])
    import json  # pylint: disable=g-import-not-at-top
    self.info_json = json.loads(blob_lib.section_text(self.info, 'info_json'))
    global LAST_LAB2D  # pylint: disable=global-statement
    LAST_LAB2D = self

  def observation_names(self) -> List[str]:
    names = []
    scalars = [n for n in self.info_json['individual_observation_names'] if n != 'RGB']
    for i in range(1, self.num_players + 1):
      names += [f'{i}.RGB', f'{i}.REWARD'] + [f'{i}.{n}' for n in scalars]
    return names + ['WORLD.RGB']


# ---------------------------------------------------------------------------------------------------------------------
# Backends
# ---------------------------------------------------------------------------------------------------------------------
class EngineBackend:
  """One env instance on the GPU through the host-buffer C-ABI calls (mp_reset_host / mp_step_host)."""

  def __init__(self, blob: bytes, seed: int, device: int = 0):
    import torch  # pylint: disable=g-import-not-at-top
    from meltingpot_b200 import engine as engine_lib  # pylint: disable=g-import-not-at-top
    self._torch = torch
    self._engine = engine_lib.Engine(blob, 1, device=device, seed=seed)
    self._host = self._engine.make_host_outputs(rgb=True, world_rgb=True, events=True)
    self._actions = torch.zeros((1, self._engine.num_players), dtype=torch.int32).pin_memory()

  def reset(self) -> None:
    self._engine.reset_host(self._host)

  def step(self, ids: Sequence[int]) -> None:
    self._actions[0] = self._torch.as_tensor(np.asarray(ids, np.int32))
    self._engine.step_host(self._actions, self._host)

  def outputs(self) -> Dict[str, np.ndarray]:
    h = self._host
    return {'rgb': h['rgb'][0].numpy().copy(), 'world_rgb': h['world_rgb'][0].numpy().copy(),
            'reward': h['reward'][0].numpy().copy(), 'scalar_obs': h['scalar_obs'][:, 0].numpy().copy(),
            'step_type': int(h['step_type'][0]), 'discount': float(h['discount'][0])}

  def events(self) -> np.ndarray:
    n = int(self._host['event_count'][0])
    return self._host['events'][0, :n].numpy().copy()

  def close(self) -> None:
    self._engine.close()


def _default_backend(blob: bytes, seed: int):
  return EngineBackend(blob, seed)


BACKEND_FACTORY: Callable[[bytes, int], Any] = _default_backend
LAST_LAB2D: Optional['Lab2d'] = None  # the most recently constructed level (diagnostics / fixture capture)


# ---------------------------------------------------------------------------------------------------------------------
# dmlab2d.Environment
# ---------------------------------------------------------------------------------------------------------------------
def _dm_env():
  from meltingpot_b200 import shims  # pylint: disable=g-import-not-at-top
  shims.install()
  import dm_env  # pylint: disable=g-import-not-at-top
  return dm_env


class Environment:
  """Stands in for `dmlab2d.Environment(env, observation_names, seed)` (a dm_env.Environment over a Lab2d level).

  Also the base class of the reference's `Lab2dWrapper` (`wrappers/base.py:26`), which overrides every method; so the
  constructor doubles as the abstract interface, like dmlab2d's own class.
  """

  def __init__(self, env: Lab2d, observation_names: Optional[Sequence[str]] = None, seed: Optional[int] = None):
    self._dm_env = _dm_env()
    self._lab = env
    self._names = list(observation_names) if observation_names is not None else env.observation_names()
    self._seed = env.env_seed if seed is None else int(seed)
    self._backend = BACKEND_FACTORY(env.blob, self._seed)
    self._P = env.num_players
    self._scalars = [n for n in env.info_json['individual_observation_names'] if n != 'RGB']
    self._out = None
    self._closed = False

  # -- dm_env API -----------------------------------------------------------------------------------------------------
  def _observation(self) -> Dict[str, np.ndarray]:
    out = self._out
    obs = {}
    for i in range(self._P):
      obs[f'{i + 1}.RGB'] = out['rgb'][i]
      obs[f'{i + 1}.REWARD'] = np.float64(out['reward'][i])
      for k, name in enumerate(self._scalars):
        obs[f'{i + 1}.{name}'] = np.float64(out['scalar_obs'][k][i])
    obs['WORLD.RGB'] = out['world_rgb']
    return {k: obs[k] for k in self._names}

  def _timestep(self):
    dm_env = self._dm_env
    self._out = self._backend.outputs()
    obs = self._observation()
    st = self._out['step_type']
    if st == 0:
      return dm_env.restart(obs)                      # reward None, discount None (multiplayer_wrapper.py:117 maps it to 0.)
    if st == 2:
      return dm_env.termination(reward=0.0, observation=obs)
    return dm_env.transition(reward=0.0, observation=obs)

  def reset(self):
    self._backend.reset()
    return self._timestep()

  def step(self, action: Mapping[str, Any]):
    space = self._lab.actions
    expected = {f'{i + 1}.{name}' for i in range(self._P) for name in space.order}
    if set(action) - expected:
      raise KeyError(f'unknown action keys {sorted(set(action) - expected)}')
    ids = []
    for i in range(self._P):
      values = [np.asarray(action.get(f'{i + 1}.{name}', default)).item() for name, (_, _, default) in zip(space.order, space.bounds)]
      ids.append(space.index(values))
    self._backend.step(ids)
    return self._timestep()

  def observation(self) -> Dict[str, np.ndarray]:
    return self._observation()

  def events(self) -> List[tuple]:
    """[(name, [b'dict', b'key', array(value), ...])] for the events:add calls on the hot path, sorted."""
    from meltingpot_b200 import engine as engine_lib  # pylint: disable=g-import-not-at-top
    out = []
    for kind, a, b in sorted(tuple(int(v) for v in row) for row in self._backend.events()):
      payload = [b'dict']
      for key, value in zip(engine_lib.EVENT_FIELDS[kind], (a, b)):
        payload += [key.encode(), np.array(float(value))]
      out.append((engine_lib.EVENT_NAMES[kind], payload))
    return out

  # -- specs ----------------------------------------------------------------------------------------------------------
  def observation_spec(self) -> Dict[str, Any]:
    specs = self._dm_env.specs
    h, w, _ = self._lab.info_json['rgb_shape']
    H, W, _ = self._lab.info_json['world_rgb_shape']
    spec = {}
    for name in self._names:
      if name == 'WORLD.RGB':
        spec[name] = specs.Array(shape=(H, W, 3), dtype=np.uint8, name=name)
      elif name.endswith('.RGB'):
        spec[name] = specs.Array(shape=(h, w, 3), dtype=np.uint8, name=name)
      else:
        spec[name] = specs.Array(shape=(), dtype=np.float64, name=name)
    return spec

  def action_spec(self) -> Dict[str, Any]:
    specs = self._dm_env.specs
    space = self._lab.actions
    return {f'{i + 1}.{name}': specs.BoundedArray(shape=(), dtype=np.int32, minimum=lo, maximum=hi, name=f'{i + 1}.{name}')
            for i in range(self._P) for name, (lo, hi, _) in zip(space.order, space.bounds)}

  def reward_spec(self):
    return self._dm_env.specs.Array(shape=(), dtype=np.float64, name='reward')

  def discount_spec(self):
    return self._dm_env.specs.BoundedArray(shape=(), dtype=np.float64, minimum=0.0, maximum=1.0, name='discount')

  # -- properties (dmlab2d's key-value debugging interface; the engine exposes none) ----------------------------------
  def list_property(self, key: str = ''):
    del key
    return []

  def read_property(self, key: str):
    raise KeyError(key)

  def write_property(self, key: str, value: str):
    raise KeyError(key)

  def close(self) -> None:
    if not self._closed:
      self._closed = True
      self._backend.close()

  def __enter__(self):
    return self

  def __exit__(self, *unused):
    self.close()


def build_modules() -> Dict[str, types.ModuleType]:
  """{module name: module} standing in for the `dmlab2d` package (registered by shims.install())."""
  mod = types.ModuleType('dmlab2d')
  mod.Lab2d = Lab2d
  mod.Environment = Environment
  runfiles = types.ModuleType('dmlab2d.runfiles_helper')
  runfiles.find = lambda: ''
  settings = types.ModuleType('dmlab2d.settings_helper')
  settings.flatten_args = flatten_args
  mod.runfiles_helper = runfiles
  mod.settings_helper = settings
  mod.__doc__ = 'B200 engine behind the dmlab2d Python surface (meltingpot_b200.lab2d_env)'
  return {'dmlab2d': mod, 'dmlab2d.runfiles_helper': runfiles, 'dmlab2d.settings_helper': settings}
