"""The reference's own Python stack over this repo's `dmlab2d` boundary module (CPU: oracle backend; needs the checkout)."""

import gzip
import json
import os
import sys
import unittest

import numpy as np
import pytest

from meltingpot_b200 import compiler, lab2d_env
from tests import ref_stack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(compiler.reference_root() is None, reason='needs the reference checkout')


def test_unflatten_inverts_flatten_and_str():
  settings = {'levelName': 'clean_up', 'numPlayers': 7, 'spriteSize': 8, 'topology': 'BOUNDED', 'flag': True, 'none': None,
              'simulation': {'map': '\nWW\nW \n', 'charPrefabMap': {'1': 'a', 'W': 'wall'}, 'rate': 0.05, 'tiny': 1e-05,
                             'prefabs': {'wall': {'name': 'wall', 'components': [{'component': 'Transform', 'kwargs': {}},
                                                                                 {'component': 'Appearance', 'kwargs': {'colors': [(1, 2, 3, 255)], 'names': ['Wall']}}]}},
                             'gameObjects': [{'name': 'avatar', 'components': [{'component': 'Avatar', 'kwargs': {'index': 1, 'neg': -1}}]}]}}
  flat = {k: str(v) for k, v in lab2d_env.flatten_args(settings).items()}
  assert flat['simulation.prefabs.wall.components.2.kwargs.colors.1.4'] == '255'
  back = lab2d_env.unflatten_args(flat)
  want = json.loads(json.dumps(settings))  # tuples -> lists
  del want['simulation']['prefabs']['wall']['components'][0]['kwargs']  # an empty dict has no flat key
  del back['simulation']['prefabs']['wall']['components'][0]
  del want['simulation']['prefabs']['wall']['components'][0]
  assert back == want
  assert isinstance(back['simulation']['charPrefabMap'], dict)  # digit keys, still a dict


@needs_reference
@pytest.mark.parametrize('name,players', [('clean_up', 7), ('territory__rooms', 9), ('coins', 2)])
def test_reference_stack_reproduces_the_committed_fixture(name, players):
  with open(os.path.join(ROOT, 'tests', 'golden', f'ref_stack_{name}.json')) as f:
    want = json.load(f)
  got = json.loads(json.dumps(ref_stack.run_reference_stack(name, players)))
  got.pop('flat_settings')
  assert got['class'] == 'meltingpot.utils.substrates.substrate.Substrate'  # the reference's class, not this repo's mirror
  assert got == want


@needs_reference
def test_compiling_from_flattened_settings_equals_compiling_from_the_config():
  # builder.py flattens the settings to "a.b.1.c" -> str for Lua (builder.py:55-67); lab2d_env un-flattens them. The blob
  # compiled from the round-tripped settings must equal the one compiled from the config's own settings.
  config = compiler.load_reference_config('clean_up')
  settings = compiler._plain(config.lab2d_settings_builder(roles=('default',) * 7, config=config))  # pylint: disable=protected-access
  flat = {k.replace('$', '.'): str(v) for k, v in lab2d_env.flatten_args(settings).items()}
  back = lab2d_env.unflatten_args(flat)
  assert compiler.compile_settings(back, config) == compiler.compile_settings(settings, config)


@needs_reference
@pytest.mark.parametrize('module', ['multiplayer_wrapper_test', 'discrete_action_wrapper_test', 'collective_reward_wrapper_test',
                                    'collective_reward_wrapper_reset_test', 'observables_wrapper_test', 'reset_wrapper_test', 'base_test'])
def test_reference_wrapper_unit_tests_pass_on_this_repos_dependency_shims(module):
  # The reference's own wrapper tests (mocks of dmlab2d.Environment) run against this repo's dmlab2d / dm_env /
  # immutabledict / reactivex stand-ins: the boundary module offers everything the wrappers touch.
  import importlib
  with ref_stack.reference_stack_on_oracle():
    mod = importlib.import_module(f'meltingpot.utils.substrates.wrappers.{module}')
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    assert suite.countTestCases() > 0
    result = unittest.TextTestRunner(stream=open(os.devnull, 'w'), verbosity=0).run(suite)
    assert result.wasSuccessful(), [str(f[1])[-600:] for f in result.failures + result.errors][:2]


@needs_reference
def test_reference_import_leaves_no_stub_modules_behind():
  compiler.load_reference_config('clean_up')
  import meltingpot  # this repo's alias package, not a stub of the checkout
  assert 'meltingpot_b200' in (meltingpot.__doc__ or '') and hasattr(meltingpot, 'substrate')
  assert not [k for k in sys.modules if k.startswith('meltingpot.configs')]


@needs_reference
@pytest.mark.parametrize('name,players', ref_stack.SUBSTRATES)
def test_reference_substrate_test_helper_accepts_the_stack(name, players):
  # The reference's own conformance helper (meltingpot/testing/substrates.py:22-68, used by substrate_test.py:24-47 for
  # every substrate): action / reward / discount / observation specs against an actual step, run on the reference's
  # stack over this repo's dmlab2d module.
  import importlib
  with ref_stack.reference_stack_on_oracle():
    helper = importlib.import_module('meltingpot.testing.substrates')
    ref_substrate = importlib.import_module('meltingpot.substrate')
    config = ref_substrate.get_config(name)
    roles = (tuple(config.default_player_roles)[0],) * players
    case = helper.SubstrateTestCase('assert_step_matches_specs')
    env = ref_substrate.build(name, roles=roles)
    try:
      case.assert_step_matches_specs(env)
      # substrate_test.py:41-47: the factory's per-player specs equal the env's
      factory = ref_substrate.get_factory(name)
      assert env.action_spec()[0] == factory.action_spec()
      assert set(env.observation_spec()[0]) >= set(factory.timestep_spec().observation)
      for key, spec in factory.timestep_spec().observation.items():
        assert env.observation_spec()[0][key] == spec, key
    finally:
      env.close()


@pytest.mark.parametrize('name,players', [('clean_up', 7), ('territory__rooms', 9)])
def test_boundary_module_compiles_builder_settings_and_refuses_to_run_without_a_gpu(name, players):
  # No reference checkout needed: the flattened settings builder.py handed to dmlab2d.Lab2d are committed. lab2d_env
  # un-flattens and compiles them; apart from the action table (full product at this boundary) the blob equals the
  # committed one. Constructing the environment needs the engine: without a CUDA device it raises, there is no CPU path.
  import torch
  from meltingpot_b200 import blob as blob_lib, engine, substrates
  with gzip.open(os.path.join(ROOT, 'tests', 'golden', f'ref_stack_settings_{name}.json.gz')) as f:
    flat = json.loads(f.read().decode())
  lab = lab2d_env.Lab2d('', flat)
  got, want = blob_lib.unpack(lab.blob), blob_lib.unpack(substrates.load_blob(name, ('default',) * players))
  for key in ('objects', 'states', 'kinds', 'comps', 'comps_f', 'init_grid', 'atlas', 'sprite_map', 'hits', 'av_table'):
    assert np.array_equal(got[key], want[key]), key
  assert got['action_table'].shape[0] == len(lab.actions.table) > want['action_table'].shape[0]
  for row in want['action_table']:  # every action of the substrate's discrete set is in the full product, at its mixed-radix index
    fields = dict(zip(['move', 'turn', 'fireZap', 'fireClean' if name == 'clean_up' else 'fireClaim'], (int(v) for v in row)))
    idx = lab.actions.index([fields[k] for k in lab.actions.order])
    assert np.array_equal(got['action_table'][idx], row)
  assert lab.observation_names()[:2] == ['1.RGB', '1.REWARD'] and lab.observation_names()[-1] == 'WORLD.RGB'
  if not torch.cuda.is_available():
    with pytest.raises(engine.EngineError, match='no CPU path'):
      lab2d_env.Environment(env=lab, observation_names=lab.observation_names(), seed=lab.env_seed)


@needs_reference
def test_flat_views_equal_the_reference_stacks_own_dmlab2d_stream():
  # The reference's innermost ObservablesWrapper emits the raw action dicts and flat TimeSteps (observables_wrapper.py:43-58).
  # meltingpot_b200.substrate.flat_action / flat_timestep (what this repo's Substrate emits on observables().dmlab2d) must
  # rebuild exactly that stream from the multiplayer-level action and TimeStep.
  import importlib
  from meltingpot_b200 import substrate as b200_substrate
  with ref_stack.reference_stack_on_oracle():
    ref_substrate = importlib.import_module('meltingpot.substrate')
    config = ref_substrate.get_config('clean_up')
    env = ref_substrate.build('clean_up', roles=('default',) * 7)
    raw_ts, raw_act = [], []
    env.observables().dmlab2d.timestep.subscribe(raw_ts.append)
    env.observables().dmlab2d.action.subscribe(raw_act.append)
    individual, global_names = list(config.individual_observation_names), list(config.global_observation_names)
    action_set = [dict(a) for a in compiler._plain(config.action_set)]  # pylint: disable=protected-access
    try:
      rng = np.random.default_rng(2)
      ts = env.reset()
      for t in range(12):
        mine = b200_substrate.flat_timestep(ts, individual, global_names)
        ref = raw_ts[-1]
        assert mine.step_type == ref.step_type and mine.reward == ref.reward and mine.discount == ref.discount
        assert set(mine.observation) <= set(ref.observation)          # (the raw env offers every observation; the wrappers select)
        for key, value in mine.observation.items():
          assert np.array_equal(value, ref.observation[key]), key
        acts = [int(a) for a in rng.integers(0, len(action_set), 7)]
        ts = env.step(acts)
        mine_act = b200_substrate.flat_action(acts, action_set)
        assert set(mine_act) == set(raw_act[-1]) and all(int(mine_act[k]) == int(raw_act[-1][k]) for k in mine_act)
    finally:
      env.close()
