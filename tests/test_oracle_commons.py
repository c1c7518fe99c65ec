"""commons_harvest__open on the CPU oracle vs what the Lua states
(/root/reference/meltingpot/lua/levels/commons_harvest/components.lua)."""

import json

import numpy as np

from meltingpot_b200 import blob as blob_lib

NOOP, FORWARD = 0, 1
N = 0
W_CELLS = 24


def _tables(blob):
  sec = blob_lib.unpack(blob)
  return sec, json.loads(blob_lib.section_text(sec, 'info_json'))


def _sprite_at(env, info, layer, cell):
  v = int(env.grid()[info['layers'].index(layer)][cell])
  return info['sprites'][(v - 1) // 4] if v else None


def test_reset_uses_inside_spawn_points_for_first_two_players(oracle, commons_blob):
  sec, info = _tables(commons_blob)
  inside = {int(c) for c in sec['spawn_cells_%d' % info['groups'].index('insideSpawnPoints')]}
  outside = {int(c) for c in sec['spawn_cells_%d' % info['groups'].index('spawnPoints')]}
  for seed in range(6):
    env = oracle.OracleEnv(commons_blob, seed)
    env.reset()
    cells = [int(y) * W_CELLS + int(x) for x, y, _, _ in env.avatars()]
    assert set(cells[:2]) == inside            # commons_harvest__open.py:517-531 (2 'Q' cells)
    assert set(cells[2:]) <= outside and len(set(cells)) == 7
    assert env.scalar_obs().shape == (7, 1)    # READY_TO_SHOOT only


def test_eaten_apple_waits_relabels_and_regrows_by_density(oracle, commons_blob):
  sec, info = _tables(commons_blob)
  apples = sec['ch_apple']
  nbr = sec['ch_nbr']
  k = int(np.argmax((nbr >= 0).sum(1)))        # an apple in the middle of a patch (12 neighbours)
  oid, cell = int(apples[k, 0]), int(apples[k, 1])
  n_nb = int((nbr[k] >= 0).sum())
  regrown = 0
  trials = 40
  for seed in range(trials):
    env = oracle.OracleEnv(commons_blob, 1000 + seed)
    env.reset()
    for p in range(7):
      env.debug_set_avatar(p, 2 + 3 * p, 12, N)  # open floor, away from the patches
    x, y = cell % W_CELLS, cell // W_CELLS
    env.debug_set_avatar(0, x, y + 1, N)
    below = (y + 1) * W_CELLS + x
    ate_below = 1.0 if _sprite_at(env, info, 'lowerPhysical', below) == 'Apple' else 0.0
    env.step(np.array([FORWARD] + [NOOP] * 6, np.int32))
    assert env.rewards()[0] == 1.0                                   # Edible:onEnter (component_library.lua:990-1002)
    assert _sprite_at(env, info, 'lowerPhysical', cell) is None     # gone the same frame
    assert _sprite_at(env, info, 'logic', cell) == 'AppleWait'
    del ate_below
    env.step(np.zeros(7, np.int32))                                  # update(): relabel to appleWait_<live neighbours>
    state = env.object_state(oid)
    names = info['kind_states'][info['kinds'].index('apple')]
    live_nb = sum(1 for j in nbr[k] if j >= 0 and env.object_state(int(apples[j, 0])) == names.index('apple'))
    assert names[state] == f'appleWait_{live_nb}' and live_nb >= n_nb - 1
    assert _sprite_at(env, info, 'background', cell) == 'Grass'      # not dessicated: it has neighbours
    # walk away; >= 3 neighbours -> regrowth probability 0.025 per frame (commons_harvest__open.py:57-58)
    env.debug_set_avatar(0, 2, 12, N)
    for _ in range(40):
      env.step(np.zeros(7, np.int32))
    regrown += env.object_state(oid) == names.index('apple')
  # P(regrow within 40 frames) = 1 - 0.975^40 = 0.64
  assert 0.35 * trials <= regrown <= 0.9 * trials


def test_isolated_apple_never_regrows_and_grass_dessicates(oracle, commons_blob):
  sec, info = _tables(commons_blob)
  apples, nbr = sec['ch_apple'], sec['ch_nbr']
  names = info['kind_states'][info['kinds'].index('apple')]
  env = oracle.OracleEnv(commons_blob, 5)
  env.reset()
  for p in range(7):
    env.debug_set_avatar(p, 2 + 3 * p, 12, N)
  # eat a whole corner patch (cells (1..3,1), (1..2,2), (1,3)): put apples to wait directly
  patch = [k for k, (oid, cell, live, g) in enumerate(apples) if cell % W_CELLS <= 3 and cell // W_CELLS <= 3]
  assert len(patch) == 6
  for k in patch:
    env.debug_set_object_state(int(apples[k, 0]), names.index('appleWait'))
  for _ in range(300):
    env.step(np.zeros(7, np.int32))
  for k in patch:
    assert names[env.object_state(int(apples[k, 0]))] == 'appleWait_0'       # no live neighbour -> probability 0
    assert _sprite_at(env, info, 'background', int(apples[k, 1])) == 'Floor'  # grass -> dessicated (:181-193)


def test_random_play_depletes_but_never_breaks_invariants(oracle, commons_blob):
  sec, info = _tables(commons_blob)
  apples, nbr = sec['ch_apple'], sec['ch_nbr']
  names = info['kind_states'][info['kinds'].index('apple')]
  env = oracle.OracleEnv(commons_blob, 77)
  env.reset()
  rng = np.random.default_rng(0)
  total = 0.0
  for t in range(500):
    env.step(rng.integers(0, 8, 7))
    total += env.rewards().sum()
    if t % 50 == 49:
      for k in range(len(apples)):
        st = names[env.object_state(int(apples[k, 0]))]
        if st.startswith('appleWait_'):
          assert 0 <= int(st.split('_')[1]) <= (nbr[k] >= 0).sum()
  assert total > 10
