"""Minimal stand-in for the `dm_env` package (used only when the real one is absent).

The reference's substrate API is expressed in dm_env terms
(`/root/reference/meltingpot/utils/substrates/specs.py:22-46`,
`/root/reference/meltingpot/utils/substrates/wrappers/multiplayer_wrapper.py:108-118`).
Neither this container nor the GPU box has dm_env installed, so the drop-in
boundary ships this small, independent re-statement of the public dm_env
surface it needs: `TimeStep`, `StepType`, `restart/transition/termination/
truncation`, `Environment` and `specs.{Array,BoundedArray,DiscreteArray}`.
If the real `dm_env` is importable it is always preferred (see `install()` in
`meltingpot_b200/shims/__init__.py`).
"""

from __future__ import annotations

import abc
import enum
import types
from typing import Any, NamedTuple

import numpy as np


class StepType(enum.IntEnum):
  FIRST = 0
  MID = 1
  LAST = 2

  def first(self) -> bool:
    return self is StepType.FIRST

  def mid(self) -> bool:
    return self is StepType.MID

  def last(self) -> bool:
    return self is StepType.LAST


class TimeStep(NamedTuple):
  step_type: Any
  reward: Any
  discount: Any
  observation: Any

  def first(self) -> bool:
    return self.step_type == StepType.FIRST

  def mid(self) -> bool:
    return self.step_type == StepType.MID

  def last(self) -> bool:
    return self.step_type == StepType.LAST


def restart(observation, reward=None, discount=None):
  # (the reference's wrapper tests pass `reward=` here: collective_reward_wrapper_reset_test.py:28-31)
  return TimeStep(StepType.FIRST, reward, discount, observation)


def transition(reward, observation, discount=1.0):
  return TimeStep(StepType.MID, reward, discount, observation)


def termination(reward, observation):
  return TimeStep(StepType.LAST, reward, 0.0, observation)


def truncation(reward, observation, discount=1.0):
  return TimeStep(StepType.LAST, reward, discount, observation)


class Environment(abc.ABC):
  """Abstract base: reset/step/specs, context manager, close()."""

  @abc.abstractmethod
  def reset(self):
    ...

  @abc.abstractmethod
  def step(self, action):
    ...

  @abc.abstractmethod
  def observation_spec(self):
    ...

  @abc.abstractmethod
  def action_spec(self):
    ...

  def reward_spec(self):
    return Array(shape=(), dtype=float, name='reward')

  def discount_spec(self):
    return BoundedArray(
        shape=(), dtype=float, minimum=0., maximum=1., name='discount')

  def close(self):
    pass

  def __enter__(self):
    return self

  def __exit__(self, exc_type, exc_value, traceback):
    del exc_type, exc_value, traceback
    self.close()


class Array:
  """Describes a numpy array or scalar shape and dtype."""
  __slots__ = ('_shape', '_dtype', '_name')

  def __init__(self, shape, dtype, name=None):
    self._shape = tuple(int(dim) for dim in shape)
    self._dtype = np.dtype(dtype)
    self._name = name

  shape = property(lambda self: self._shape)
  dtype = property(lambda self: self._dtype)
  name = property(lambda self: self._name)

  def __repr__(self):
    return 'Array(shape={}, dtype={}, name={})'.format(
        self.shape, repr(self.dtype), repr(self.name))

  def __eq__(self, other):
    # As in dm_env, equality ignores `name`.
    if not isinstance(other, Array):
      return False
    return (type(self) is type(other) and self.shape == other.shape and
            self.dtype == other.dtype)

  def __ne__(self, other):
    return not self == other

  __hash__ = None

  def _fail_validation(self, message, *args):
    message %= args
    if self.name:
      message += ' for spec %s' % self.name
    raise ValueError(message)

  def validate(self, value):
    value = np.asarray(value)
    if value.shape != self.shape:
      self._fail_validation('Expected shape %r but found %r', self.shape,
                            value.shape)
    if value.dtype != self.dtype:
      self._fail_validation('Expected dtype %s but found %s', self.dtype,
                            value.dtype)
    return value

  def generate_value(self):
    return np.zeros(shape=self.shape, dtype=self.dtype)

  def _get_constructor_kwargs(self):
    return dict(shape=self._shape, dtype=self._dtype, name=self._name)

  def replace(self, **kwargs):
    all_kwargs = self._get_constructor_kwargs()
    all_kwargs.update(kwargs)
    return type(self)(**all_kwargs)

  def __reduce__(self):
    return (_rebuild, (type(self), self._get_constructor_kwargs()))


def _rebuild(cls, kwargs):
  return cls(**kwargs)


class BoundedArray(Array):
  """An `Array` spec with inclusive minimum and maximum."""
  __slots__ = ('_minimum', '_maximum')

  def __init__(self, shape, dtype, minimum, maximum, name=None):
    super().__init__(shape, dtype, name)
    try:
      bcast_minimum = np.broadcast_to(minimum, shape=shape)
      bcast_maximum = np.broadcast_to(maximum, shape=shape)
    except ValueError as e:
      raise ValueError('minimum/maximum not compatible with shape') from e
    if np.any(bcast_minimum > bcast_maximum):
      raise ValueError('All values in `minimum` must be <= `maximum`.')
    self._minimum = np.array(minimum, dtype=self.dtype)
    self._minimum.setflags(write=False)
    self._maximum = np.array(maximum, dtype=self.dtype)
    self._maximum.setflags(write=False)

  minimum = property(lambda self: self._minimum)
  maximum = property(lambda self: self._maximum)

  def __repr__(self):
    return ('BoundedArray(shape={}, dtype={}, name={}, minimum={}, maximum={})'
            .format(self.shape, repr(self.dtype), repr(self.name),
                    self._minimum, self._maximum))

  def __eq__(self, other):
    if not isinstance(other, BoundedArray):
      return False
    return (super().__eq__(other) and
            (self.minimum == other.minimum).all() and
            (self.maximum == other.maximum).all())

  __hash__ = None

  def validate(self, value):
    value = np.asarray(value)
    super().validate(value)
    if (value < self.minimum).any() or (value > self.maximum).any():
      self._fail_validation(
          'Values were not all within bounds %s <= %s <= %s', self.minimum,
          value, self.maximum)
    return value

  def generate_value(self):
    return (np.ones(shape=self.shape, dtype=self.dtype) *
            self.dtype.type(self.minimum))

  def _get_constructor_kwargs(self):
    kwargs = super()._get_constructor_kwargs()
    kwargs.update(minimum=self._minimum, maximum=self._maximum)
    return kwargs


class DiscreteArray(BoundedArray):
  """Represents a discretely valued scalar in [0, num_values)."""
  __slots__ = ('_num_values',)

  def __init__(self, num_values, dtype=np.int32, name=None):
    if num_values <= 0 or not np.issubdtype(type(num_values), np.integer):
      raise ValueError('`num_values` must be a positive integer, got {}.'
                       .format(num_values))
    if not np.issubdtype(dtype, np.integer):
      raise ValueError('`dtype` must be integer, got {}.'.format(dtype))
    super().__init__(
        shape=(), dtype=dtype, minimum=0, maximum=num_values - 1, name=name)
    self._num_values = int(num_values)

  num_values = property(lambda self: self._num_values)

  def __repr__(self):
    return 'DiscreteArray(shape={}, dtype={}, name={}, minimum={}, ' \
           'maximum={}, num_values={})'.format(
               self.shape, repr(self.dtype), repr(self.name), self.minimum,
               self.maximum, self.num_values)

  def _get_constructor_kwargs(self):
    return dict(num_values=self._num_values, dtype=self._dtype,
                name=self._name)


def build_modules():
  """Returns (dm_env, dm_env.specs) module objects backed by this file."""
  dm_env = types.ModuleType('dm_env')
  specs = types.ModuleType('dm_env.specs')
  for cls in (Array, BoundedArray, DiscreteArray):
    setattr(specs, cls.__name__, cls)
  for obj in (StepType, TimeStep, Environment):
    setattr(dm_env, obj.__name__, obj)
  for fn in (restart, transition, termination, truncation):
    setattr(dm_env, fn.__name__, fn)
  dm_env.specs = specs
  dm_env.__meltingpot_b200_shim__ = True
  return dm_env, specs
