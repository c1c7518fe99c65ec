"""Public API: drop-in for `meltingpot.substrate` backed by the B200 engine.

Mirrors `/root/reference/meltingpot/substrate.py:38-113`:
  SUBSTRATES, get_config(name), build(name, *, roles), build_from_config(config, *, roles),
  get_factory(name), get_factory_from_config(config)
and adds `build_batched(...)` for thousands of env instances as torch tensors.

`build(...)` returns a `Substrate` with the reference's dm_env surface and timestep layout
(`/root/reference/meltingpot/utils/substrates/substrate.py:47-104`, wrapper stack `:107-139`):
  * `step(actions)` takes one discrete action id per player (DiscreteActionWrapper);
  * `timestep.reward` is a list of P float64 scalars, `discount` 0.0 on FIRST/LAST else 1.0
    (MultiplayerWrapper, multiplayer_wrapper.py:108-118);
  * `timestep.observation` is a list of P dicts with the configured per-player and global
    observations plus `COLLECTIVE_REWARD` (collective_reward_wrapper.py:39-50);
  * a step after LAST starts a new episode and returns FIRST.
All state transition and rendering runs in the CUDA engine; there is no CPU path.
"""

from __future__ import annotations

import dataclasses
import json
from typing import Any, Callable, Collection, Dict, List, Mapping, Optional, Sequence

import numpy as np

from meltingpot_b200 import blob as blob_lib
from meltingpot_b200 import shims
from meltingpot_b200 import specs as specs_lib
from meltingpot_b200 import substrates as substrate_blobs

shims.install()
import dm_env  # noqa: E402  pylint: disable=g-import-not-at-top,g-bad-import-order
from ml_collections import config_dict  # noqa: E402  pylint: disable=g-import-not-at-top,g-bad-import-order

SUBSTRATES = frozenset(substrate_blobs.PRECOMPILED)
_COLLECTIVE_REWARD_OBS = 'COLLECTIVE_REWARD'
_SCALAR_NAMES = {0: 'READY_TO_SHOOT', 1: 'NUM_OTHERS_WHO_CLEANED_THIS_STEP',
                 2: 'MISMATCHED_COIN_COLLECTED_BY_PARTNER'}  # include/mpb_format.h MpbScalarObs
_MAX_SEED = 2**32 - 1


# ---------------------------------------------------------------------------------------------
# Config
# ---------------------------------------------------------------------------------------------
def _info(blob: bytes) -> Dict[str, Any]:
  return json.loads(blob_lib.section_text(blob_lib.unpack(blob), 'info_json'))


def _config_from_info(name: str, info: Mapping[str, Any]) -> config_dict.ConfigDict:
  config = config_dict.ConfigDict()
  config.substrate_name = name
  config.action_set = tuple(dict(a) for a in info['action_set'])
  config.individual_observation_names = list(info['individual_observation_names'])
  config.global_observation_names = list(info['global_observation_names'])
  config.action_spec = specs_lib.action(len(info['action_set']))
  obs = {}
  for key in info['individual_observation_names']:
    obs[key] = specs_lib.rgb(*info['rgb_shape'][:2]) if key == 'RGB' else specs_lib.float64()
  for key in info['global_observation_names']:
    obs[key] = specs_lib.rgb(*info['world_rgb_shape'][:2])
  config.timestep_spec = specs_lib.timestep(obs)
  config.valid_roles = frozenset(info['valid_roles'])
  config.default_player_roles = tuple(info['default_player_roles'])
  return config


def get_config(name: str) -> config_dict.ConfigDict:
  """Returns the locked configuration for the specified substrate (substrate.py:41-54)."""
  if name not in SUBSTRATES:
    raise ValueError(f'{name} not in {sorted(SUBSTRATES)}.')
  blob = substrate_blobs.load_blob(name)
  return _config_from_info(name, _info(blob)).lock()


# ---------------------------------------------------------------------------------------------
# Observables (reactivex is not available here; this is the small subset Substrate exposes)
# ---------------------------------------------------------------------------------------------
class Subject:
  """Minimal hot observable: subscribe(on_next, on_error, on_completed)."""

  def __init__(self):
    self._observers: List[Any] = []

  def subscribe(self, on_next=None, on_error=None, on_completed=None):
    if on_next is not None and not callable(on_next):  # observer object
      observer = on_next
      entry = (getattr(observer, 'on_next', None), getattr(observer, 'on_error', None),
               getattr(observer, 'on_completed', None))
    else:
      entry = (on_next, on_error, on_completed)
    self._observers.append(entry)
    return entry

  def on_next(self, value):
    for fn, _, _ in list(self._observers):
      if fn:
        fn(value)

  def on_completed(self):
    for _, _, fn in list(self._observers):
      if fn:
        fn()
    self._observers.clear()


@dataclasses.dataclass(frozen=True)
class Lab2dObservables:
  """Same fields as the reference's Lab2dObservables (wrappers/observables.py:32-45): the raw dmlab2d-level stream."""
  action: Subject      # {"1.move": ..., "1.turn": ..., ...} per step
  timestep: Subject    # dm_env.TimeStep with the flat {"1.RGB", "1.REWARD", ..., "WORLD.RGB"} observation dict
  events: Subject


@dataclasses.dataclass(frozen=True)
class SubstrateObservables:
  """Same fields as the reference's SubstrateObservables (substrate.py:32-44)."""
  action: Subject
  timestep: Subject
  events: Subject
  dmlab2d: Optional[Lab2dObservables] = None


def flat_action(action: Sequence[int], action_set: Sequence[Mapping[str, int]]) -> Dict[str, np.ndarray]:
  """What the wrapper stack would hand to dmlab2d for these discrete actions (discrete_action_wrapper.py:97-100,
  multiplayer_wrapper.py:120-130): {"<player>.<field>": int32 scalar}."""
  out = {}
  for i, a in enumerate(action):
    for key, value in action_set[int(a)].items():
      out[f'{i + 1}.{key}'] = np.array(value, dtype=np.int32)
  return out


def flat_timestep(timestep: 'dm_env.TimeStep', individual: Sequence[str], global_names: Sequence[str]) -> 'dm_env.TimeStep':
  """The dmlab2d-level view of a multiplayer TimeStep: flat observation dict with "{i}.REWARD" entries, reward None on
  FIRST else 0.0, discount None on FIRST (the inverse of multiplayer_wrapper.py:80-118)."""
  obs = {}
  for i, (player, reward) in enumerate(zip(timestep.observation, timestep.reward)):
    for name in individual:
      obs[f'{i + 1}.{name}'] = player[name]
    obs[f'{i + 1}.REWARD'] = np.float64(reward)
  for name in global_names:
    obs[name] = timestep.observation[0][name]
  first = timestep.step_type == dm_env.StepType.FIRST
  return dm_env.TimeStep(step_type=timestep.step_type, reward=None if first else 0.0,
                         discount=None if first else timestep.discount, observation=obs)


# ---------------------------------------------------------------------------------------------
# Batched substrate (tensors)
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class BatchedTimeStep:
  """One timestep of B env instances; all fields are CUDA tensors viewing engine buffers."""
  step_type: Any   # int64 [B]
  reward: Any      # float64 [B, P]
  discount: Any    # float64 [B]
  observation: Dict[str, Any]


class BatchedSubstrate:
  """`num_envs` independent instances of a substrate on one GPU.

  Observations are zero-copy views of the engine's output buffers and are overwritten by the
  next `step`/`reset`; clone what must be kept.
  """

  def __init__(self, blob: bytes, num_envs: int, device: int = 0, seed: Optional[int] = None,
               env_index_base: int = 0, world_rgb: bool = True):
    from meltingpot_b200 import engine as engine_lib  # pylint: disable=g-import-not-at-top
    if seed is None:
      seed = int(np.random.randint(1, _MAX_SEED))
    self._info = _info(blob)
    flags = engine_lib.MP_FLAG_RENDER_PLAYERS | (engine_lib.MP_FLAG_RENDER_WORLD if world_rgb else 0)
    self._engine = engine_lib.Engine(blob, num_envs, device=device, seed=seed,
                                     env_index_base=env_index_base, flags=flags)
    self.num_envs = num_envs
    self.num_players = self._engine.num_players
    self.num_actions = self._engine.num_actions
    self.seed = seed
    sections = blob_lib.unpack(blob)
    self._scalar_names = [_SCALAR_NAMES[int(k)] for k in sections['scalar_obs']]
    self._world_rgb = world_rgb

  @property
  def engine(self):
    return self._engine

  def _timestep(self) -> BatchedTimeStep:
    e = self._engine
    obs = {'RGB': e.rgb}
    for k, name in enumerate(self._scalar_names):
      obs[name] = e.scalar_obs[k]
    if self._world_rgb:
      obs['WORLD.RGB'] = e.world_rgb
    obs[_COLLECTIVE_REWARD_OBS] = e.reward.sum(dim=1)
    return BatchedTimeStep(step_type=e.step_type, reward=e.reward, discount=e.discount, observation=obs)

  def reset(self, mask=None) -> BatchedTimeStep:
    self._engine.reset(mask)
    return self._timestep()

  def step(self, actions) -> BatchedTimeStep:
    """actions: integer tensor [B, P] on the engine's device (int32 preferred)."""
    import torch  # pylint: disable=g-import-not-at-top
    if actions.dtype != torch.int32:
      actions = actions.to(torch.int32)
    self._engine.step(actions.contiguous())
    return self._timestep()

  def debug_observations(self, layer: bool = True, zap_matrix: bool = True):
    """The reference's debug observations of the current timestep as int32 tensors: 'POSITION' [B, P, 2],
    'ORIENTATION' [B, P] (specs.py:39-44), 'LAYER' [B, P, view_h, view_w, layers], 'ZAP_MATRIX' [B, P, P]."""
    return self._engine.debug_observations(layer=layer, zap_matrix=zap_matrix)

  def events(self):
    """(event_count int32 [B], events int32 [B, max_events, 3]) of the last step; rows are (type, a, b), unordered."""
    return self._engine.event_count, self._engine.events

  def save_state(self) -> bytes:
    """Snapshot of every env instance (no reference counterpart)."""
    return self._engine.save_state()

  def load_state(self, snapshot: bytes) -> BatchedTimeStep:
    """Restores a snapshot taken from an identically built BatchedSubstrate; returns the timestep it held."""
    self._engine.load_state(snapshot)
    return self._timestep()

  def action_spec(self):
    return tuple(specs_lib.action(self.num_actions) for _ in range(self.num_players))

  def close(self):
    self._engine.close()

  def __enter__(self):
    return self

  def __exit__(self, *unused):
    self.close()


# ---------------------------------------------------------------------------------------------
# dm_env substrate (one env, numpy)
# ---------------------------------------------------------------------------------------------
class Substrate(dm_env.Environment):
  """dm_env view of a single env instance (the reference's `Substrate`)."""

  def __init__(self, blob: bytes, config: config_dict.ConfigDict, device: int = 0,
               env_seed: Optional[int] = None):
    import torch  # pylint: disable=g-import-not-at-top
    self._torch = torch
    self._config = config
    self._batched = BatchedSubstrate(blob, 1, device=device, seed=env_seed)
    self._num_players = self._batched.num_players
    self._individual = list(config.individual_observation_names)
    self._global = list(config.global_observation_names)
    self._action_subject, self._timestep_subject, self._events_subject = Subject(), Subject(), Subject()
    self._raw = Lab2dObservables(action=Subject(), timestep=Subject(), events=Subject())
    self._observables = SubstrateObservables(dmlab2d=self._raw, action=self._action_subject, timestep=self._timestep_subject,
                                             events=self._events_subject)
    self._closed = False
    self._last_observation = None
    self._last_events = np.zeros((0, 3), np.int32)
    self._host = None

  # -- helpers --------------------------------------------------------------------------------
  def _host_buffers(self):
    """Pinned host buffers of the B = 1 view: one host-buffer C-ABI call per step fills them (three device->host
    copies: images, WORLD.RGB, and the packed scalar block; plus the events), then the stream is synchronised once."""
    if self._host is None:
      eng = self._batched.engine
      self._host = eng.make_host_outputs(rgb=True, world_rgb=self._batched._world_rgb, events=True)  # pylint: disable=protected-access
      self._host_actions = eng.make_host_actions()
      self._scalar_index = {name: k for k, name in enumerate(self._batched._scalar_names)}  # pylint: disable=protected-access
    return self._host

  def _to_timestep(self) -> dm_env.TimeStep:
    host = self._host
    step_type = dm_env.StepType(int(host['step_type'][0]))
    rewards = [np.float64(r) for r in host['reward'][0].numpy()]
    # fresh arrays every step, as the reference returns (the pinned buffers are overwritten by the next call)
    rgb = host['rgb'][0].numpy().copy()
    shared = {}
    for name in self._global:
      shared[name] = host['world_rgb'][0].numpy().copy()
    collective = np.sum(rewards)
    observations = []
    for i in range(self._num_players):
      obs = {_COLLECTIVE_REWARD_OBS: collective}
      for name in self._individual:
        obs[name] = rgb[i] if name == 'RGB' else np.float64(host['scalar_obs'][self._scalar_index[name], 0, i])
      for name in self._global:
        obs[name] = shared[name]  # the same array object in every player's dict
      observations.append(obs)
    self._last_observation = observations
    n = int(host['event_count'][0])
    if n > host['events'].shape[1]:
      raise RuntimeError(f'{n} events in one step exceed the engine\'s max_events {host["events"].shape[1]}')
    self._last_events = host['events'][0, :n].numpy().copy()
    return dm_env.TimeStep(step_type=step_type, reward=rewards, discount=float(host['discount'][0]),
                           observation=observations)

  # -- dm_env API -------------------------------------------------------------------------------
  def _emit_raw(self, timestep, action=None) -> None:
    """observables().dmlab2d: the dmlab2d-level stream the reference's innermost ObservablesWrapper emits
    (observables_wrapper.py:43-58), built only while somebody is subscribed."""
    raw = self._raw
    if action is not None and raw.action._observers:  # pylint: disable=protected-access
      raw.action.on_next(flat_action(action, self._config.action_set))
    if raw.timestep._observers:  # pylint: disable=protected-access
      raw.timestep.on_next(flat_timestep(timestep, self._individual, self._global))
    if raw.events._observers:  # pylint: disable=protected-access
      for event in self.events():
        raw.events.on_next(event)

  def reset(self) -> dm_env.TimeStep:
    self._batched.engine.reset_host(self._host_buffers())
    timestep = self._to_timestep()
    self._emit_raw(timestep)
    self._timestep_subject.on_next(timestep)
    for event in self.events():
      self._events_subject.on_next(event)
    return timestep

  def step(self, action: Sequence[int]) -> dm_env.TimeStep:
    if len(action) != self._num_players:
      raise ValueError(f'expected {self._num_players} actions, got {len(action)}')
    specs = self.action_spec()
    for a, spec in zip(action, specs):
      if not 0 <= int(a) < spec.num_values:
        raise ValueError(f'action {a} out of range [0, {spec.num_values})')
    self._action_subject.on_next(action)
    host = self._host_buffers()
    self._host_actions[0] = self._torch.as_tensor(np.asarray(action, np.int32))
    self._batched.engine.step_host(self._host_actions, host)
    timestep = self._to_timestep()
    self._emit_raw(timestep, action)
    self._timestep_subject.on_next(timestep)
    for event in self.events():
      self._events_subject.on_next(event)
    return timestep

  def observation(self) -> Sequence[Mapping[str, np.ndarray]]:
    return self._last_observation

  def events(self) -> Sequence[tuple]:
    """Events of the last reset/step in dmlab2d's shape: (name, [b'dict', b'key', array(value), ...]).

    Covers the events:add calls on the hot path (include/mp_engine.h, mp_buffers.events). dmlab2d's own
    order within a step is engine-defined and unpinned; here they are sorted by (type, arguments).
    """
    from meltingpot_b200 import engine as engine_lib  # pylint: disable=g-import-not-at-top
    rows = sorted(tuple(int(v) for v in row) for row in self._last_events)
    out = []
    for kind, a, b in rows:
      payload = [b'dict']
      for key, value in zip(engine_lib.EVENT_FIELDS[kind], (a, b)):
        payload += [key.encode(), np.array(float(value))]
      out.append((engine_lib.EVENT_NAMES[kind], payload))
    return out

  def action_spec(self) -> Sequence['dm_env.specs.DiscreteArray']:
    return tuple(self._config.action_spec for _ in range(self._num_players))

  def observation_spec(self) -> Sequence[Mapping[str, 'dm_env.specs.Array']]:
    spec = dict(self._config.timestep_spec.observation)
    spec[_COLLECTIVE_REWARD_OBS] = dm_env.specs.Array(shape=(), dtype=np.float64, name=_COLLECTIVE_REWARD_OBS)
    return tuple(dict(spec) for _ in range(self._num_players))

  def reward_spec(self) -> Sequence['dm_env.specs.Array']:
    return tuple(self._config.timestep_spec.reward for _ in range(self._num_players))

  def discount_spec(self):
    return self._config.timestep_spec.discount

  def observables(self) -> SubstrateObservables:
    return self._observables

  # dmlab2d's key-value debugging interface (wrappers/base.py:66-80): the engine exposes no properties.
  def list_property(self, key: str = ''):
    del key
    return []

  def read_property(self, key: str):
    raise KeyError(key)

  def write_property(self, key: str, value: str):
    del value
    raise KeyError(key)

  def close(self) -> None:
    if not self._closed:
      self._closed = True
      self._batched.close()
      for subject in (self._raw.action, self._raw.timestep, self._raw.events):
        subject.on_completed()
      self._action_subject.on_completed()
      self._timestep_subject.on_completed()
      self._events_subject.on_completed()


# ---------------------------------------------------------------------------------------------
# Factories
# ---------------------------------------------------------------------------------------------
def _validate_roles(config, roles: Sequence[str]) -> None:
  invalid = set(roles) - set(config.valid_roles)
  if invalid:  # configs/substrates/__init__.py:42-45
    raise ValueError(f'Invalid roles: {invalid!r}. Must be one of {config.valid_roles!r}')


class SubstrateFactory:
  """Mirrors `/root/reference/meltingpot/utils/substrates/substrate_factory.py:24-95`."""

  def __init__(self, name: str, config: config_dict.ConfigDict, device: int = 0):
    self._name = name
    self._config = config
    self._device = device

  def valid_roles(self) -> Collection[str]:
    return frozenset(self._config.valid_roles)

  def default_player_roles(self) -> Sequence[str]:
    return tuple(self._config.default_player_roles)

  def timestep_spec(self):
    return self._config.timestep_spec

  def action_spec(self):
    return self._config.action_spec

  def build(self, roles: Sequence[str], env_seed: Optional[int] = None) -> Substrate:
    _validate_roles(self._config, roles)
    blob = substrate_blobs.load_blob(self._name, tuple(roles))
    return Substrate(blob, self._config, device=self._device, env_seed=env_seed)

  def build_batched(self, roles: Sequence[str], num_envs: int, seed: Optional[int] = None,
                    env_index_base: int = 0, world_rgb: bool = True) -> BatchedSubstrate:
    _validate_roles(self._config, roles)
    blob = substrate_blobs.load_blob(self._name, tuple(roles))
    return BatchedSubstrate(blob, num_envs, device=self._device, seed=seed,
                            env_index_base=env_index_base, world_rgb=world_rgb)


def get_factory(name: str, device: int = 0) -> SubstrateFactory:
  return SubstrateFactory(name, get_config(name), device=device)


def get_factory_from_config(config: config_dict.ConfigDict, device: int = 0) -> SubstrateFactory:
  return SubstrateFactory(config.substrate_name, config, device=device)


def build(name: str, *, roles: Sequence[str], env_seed: Optional[int] = None, device: int = 0) -> Substrate:
  """Builds an instance of the specified substrate (substrate.py:57-70)."""
  return get_factory(name, device).build(roles, env_seed=env_seed)


def build_from_config(config: config_dict.ConfigDict, *, roles: Sequence[str],
                      env_seed: Optional[int] = None, device: int = 0) -> Substrate:
  return get_factory_from_config(config, device).build(roles, env_seed=env_seed)


def build_batched(name: str, *, roles: Sequence[str], num_envs: int, device: int = 0,
                  seed: Optional[int] = None, env_index_base: int = 0,
                  world_rgb: bool = True) -> BatchedSubstrate:
  """Builds `num_envs` instances on one GPU; see `BatchedSubstrate`."""
  return get_factory(name, device).build_batched(roles, num_envs, seed=seed,
                                                 env_index_base=env_index_base, world_rgb=world_rgb)
