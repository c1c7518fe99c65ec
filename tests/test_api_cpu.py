"""Host-side logic that must hold without a GPU: shims, config/spec surface, C ABI symbols."""

import ctypes
import os
import re

import numpy as np
import pytest

from meltingpot_b200 import engine, shims, specs, substrate

shims.install()
import dm_env  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dm_env_surface():
  ts = dm_env.restart({'a': 1})
  assert ts.first() and ts.reward is None and ts.discount is None
  assert dm_env.transition(1.0, {}).mid() and dm_env.transition(1.0, {}).discount == 1.0
  last = dm_env.termination(2.0, {})
  assert last.last() and last.discount == 0.0
  assert int(dm_env.StepType.FIRST) == 0 and int(dm_env.StepType.LAST) == 2


def test_spec_equality_ignores_name_like_dm_env():
  # substrate_test.py:41-47 compares per-player specs with `==`; names differ, shapes do not.
  a = dm_env.specs.Array((88, 88, 3), np.uint8, name='RGB')
  b = dm_env.specs.Array((88, 88, 3), np.uint8, name='1.RGB')
  assert a == b and a != dm_env.specs.Array((88, 88, 3), np.int32)
  d = dm_env.specs.DiscreteArray(9, dtype=np.int64, name='action')
  assert d.num_values == 9 and d.maximum == 8 and d.replace(num_values=8).num_values == 8
  d.validate(np.int64(3))
  with pytest.raises(ValueError):
    d.validate(np.int64(9))
  with pytest.raises(ValueError):
    a.validate(np.zeros((88, 88, 3), np.float32))


def test_config_dict_lock():
  from ml_collections import config_dict
  c = config_dict.ConfigDict()
  c.x = 1
  c.lock()
  with pytest.raises(AttributeError):
    c.y = 2
  with c.unlocked():
    c.y = 2
  assert c.y == 2 and c.is_locked and c.to_dict() == {'x': 1, 'y': 2}


def test_clean_up_config_matches_reference_api():
  config = substrate.get_config('clean_up')
  # clean_up.py:461-483: ids map to NOOP, FORWARD, BACKWARD, STEP_LEFT, STEP_RIGHT, TURN_LEFT, TURN_RIGHT, ZAP, CLEAN
  assert len(config.action_set) == 9
  assert config.action_set[2] == {'move': 3, 'turn': 0, 'fireZap': 0, 'fireClean': 0}
  assert config.action_set[4]['move'] == 2 and config.action_set[5]['turn'] == -1
  assert config.action_set[7]['fireZap'] == 1 and config.action_set[8]['fireClean'] == 1
  assert list(config.individual_observation_names) == ['RGB', 'READY_TO_SHOOT', 'NUM_OTHERS_WHO_CLEANED_THIS_STEP']
  assert list(config.global_observation_names) == ['WORLD.RGB']
  obs = config.timestep_spec.observation
  assert obs['RGB'] == specs.rgb(88, 88) and obs['WORLD.RGB'] == specs.rgb(168, 240)
  assert obs['READY_TO_SHOOT'].dtype == np.float64 and obs['READY_TO_SHOOT'].shape == ()
  assert config.action_spec.num_values == 9 and config.action_spec.dtype == np.int64
  assert config.valid_roles == frozenset({'default'}) and config.default_player_roles == ('default',) * 7
  assert config.timestep_spec.reward.dtype == np.float64
  with pytest.raises(ValueError):
    substrate.get_config('not_a_substrate')


def test_invalid_roles_raise_value_error():
  # configs/substrates/__init__.py:42-45
  with pytest.raises(ValueError, match='Invalid roles'):
    substrate.build('clean_up', roles=('default', 'cleaner'))


def test_c_abi_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, 'include', 'mp_engine.h')).read()
  declared = set(re.findall(r'\b(mp_[a-z_]+)\s*\(', header))
  declared.discard('mp_engine')
  assert declared == set(engine.EXPORTED_SYMBOLS), declared ^ set(engine.EXPORTED_SYMBOLS)
  lib = engine.load_library()
  for sym in declared:
    assert hasattr(lib, sym), sym
  assert b'sm_100a' in lib.mp_version()


def _cuda_available():
  import torch
  return torch.cuda.is_available()


@pytest.mark.skipif(_cuda_available(), reason='checks the no-GPU failure mode')
def test_engine_fails_loudly_without_gpu(clean_up_blob):
  lib = engine.load_library()
  handle = ctypes.c_void_p()
  rc = lib.mp_create(clean_up_blob, len(clean_up_blob), 4, 0, ctypes.c_uint64(1), ctypes.c_uint64(0),
                     ctypes.c_uint32(3), ctypes.byref(handle))
  assert rc == -4 and not handle.value  # MP_E_NO_DEVICE: there is no CPU path to fall back to
  assert b'no CPU path' in lib.mp_last_error()
  with pytest.raises(engine.EngineError):
    engine.Engine(clean_up_blob, 4)
  with pytest.raises(engine.EngineError):
    substrate.build('clean_up', roles=('default',) * 7)


def test_product_package_never_imports_the_oracle():
  pkg = os.path.join(ROOT, 'meltingpot_b200')
  bad = re.compile(r'^\s*(#\s*include\s*[<"][^>"]*oracle|from\s+oracle\b|import\s+oracle\b)|liboracle|CDLL\([^)]*oracle',
                   re.MULTILINE)
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        text = open(os.path.join(dirpath, f), errors='ignore').read()
        assert not bad.search(text), f
