"""The CPU oracle against what the reference's Lua states (and the engine policy ledger).

Each test cites the reference lines it is derived from. These pin the oracle (and hence, through
the -m gpu parity tests, the CUDA engine) to the reference's documented behaviour; bit-level parity
with a real DMLab2D run remains unpinned.
"""

import json

import numpy as np
import pytest

from meltingpot_b200 import blob as blob_lib

NOOP, FORWARD, BACKWARD, STEP_LEFT, STEP_RIGHT, TURN_LEFT, TURN_RIGHT, ZAP, CLEAN = range(9)
N, E, S, W = range(4)
P = 7
W_CELLS = 30


@pytest.fixture(scope='module')
def tables(clean_up_blob):
  sec = blob_lib.unpack(clean_up_blob)
  info = json.loads(blob_lib.section_text(sec, 'info_json'))
  return sec, info


def fresh(oracle, blob, seed=1, park=True):
  env = oracle.OracleEnv(blob, seed)
  env.reset()
  if park:  # line the avatars up on the grass, out of each other's way
    for p in range(P):
      env.debug_set_avatar(p, 2 + 3 * p, 18, N)
  return env


def act(**kw):
  a = [NOOP] * P
  for k, v in kw.items():
    a[int(k[1:])] = v
  return np.array(a, np.int32)


def layer_sprites(env, info, layer_name):
  grid = env.grid()[info['layers'].index(layer_name)]
  return {int(c): info['sprites'][(int(v) - 1) // 4] for c, v in enumerate(grid) if v}


def test_philox_known_answers(oracle):
  # Random123 kat_vectors for philox4x32-10.
  assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
  assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
  assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [
      0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_reset_timestep(oracle, clean_up_blob):
  env = oracle.OracleEnv(clean_up_blob, 3)
  assert env.reset() == 0 and env.step_type() == 0
  assert env.discount() == 0.0 and (env.rewards() == 0).all()        # multiplayer_wrapper.py:116-117
  assert (env.scalar_obs()[:, 0] == 1.0).all() and (env.scalar_obs()[:, 1] == 0.0).all()
  av = env.avatars()
  assert (av[:, 3] == 1).all() and len({(x, y) for x, y, _, _ in av}) == P  # spawn without replacement
  assert env.counters()['dirt'] == 79 and env.counters()['clean'] == 68     # clean_up.py map: F / H cells


def test_relative_moves_and_turns(oracle, clean_up_blob):
  env = fresh(oracle, clean_up_blob)
  env.debug_set_avatar(0, 5, 10, S)
  env.step(act(p0=STEP_RIGHT))            # facing S, moveRel('E') moves x-1 (game_object_test.lua:281-293)
  assert tuple(env.avatars()[0][:3]) == (4, 10, S)
  env.step(act(p0=FORWARD))
  assert tuple(env.avatars()[0][:3]) == (4, 11, S)
  env.step(act(p0=BACKWARD))
  assert tuple(env.avatars()[0][:3]) == (4, 10, S)
  env.step(act(p0=TURN_LEFT))             # turn(3) from S -> E (game_object_test.lua:347-362)
  assert env.avatars()[0][2] == E
  env.step(act(p0=TURN_RIGHT))
  assert env.avatars()[0][2] == S
  env.step(act(p0=STEP_LEFT))             # moveRel('W') when facing S moves x+1
  assert tuple(env.avatars()[0][:2]) == (5, 10)


def test_walls_and_avatars_block(oracle, clean_up_blob):
  env = fresh(oracle, clean_up_blob)
  env.debug_set_avatar(0, 1, 10, W)
  env.step(act(p0=FORWARD))               # wall at x = 0 (superOverlay, same layer as avatars; clean_up.py:297,644)
  assert tuple(env.avatars()[0][:2]) == (1, 10)
  env.debug_set_avatar(0, 5, 10, E)
  env.debug_set_avatar(1, 6, 10, W)
  env.step(act(p0=FORWARD, p1=FORWARD))   # swapping places is impossible: both targets are occupied
  assert tuple(env.avatars()[0][:2]) == (5, 10) and tuple(env.avatars()[1][:2]) == (6, 10)


def test_contested_cell_goes_to_exactly_one_avatar_in_random_order(oracle, clean_up_blob):
  winners = set()
  for seed in range(24):
    env = fresh(oracle, clean_up_blob, seed=seed)
    env.debug_set_avatar(0, 5, 10, E)
    env.debug_set_avatar(1, 7, 10, W)
    env.step(act(p0=FORWARD, p1=FORWARD))
    pos = [tuple(a[:2]) for a in env.avatars()[:2]]
    assert sorted(pos) in ([(5, 10), (6, 10)], [(6, 10), (7, 10)])
    winners.add(0 if pos[0] == (6, 10) else 1)
  assert winners == {0, 1}  # policy A.7: per-frame random order, not index order


def test_zap_removes_and_respawns_after_exactly_50_frames(oracle, clean_up_blob, tables):
  _, info = tables
  env = fresh(oracle, clean_up_blob)
  env.debug_set_avatar(0, 5, 10, E)
  env.debug_set_avatar(1, 6, 10, W)
  env.debug_set_avatar(2, 7, 10, W)
  env.step(act(p0=ZAP))
  assert ('zap', 1, 2) in env.events() and ('zap', 1, 3) not in env.events()  # beam stops at the first avatar
  alive = env.avatars()[:, 3]
  assert alive[1] == 0 and alive[2] == 1 and alive[0] == 1               # removeHitPlayer (avatar_library.lua:666-668)
  assert (env.rewards() == 0).all()                                       # penalty / reward are 0 in clean_up
  assert (env.rgb()[1] == 80).all()                                       # policy A.13: out-of-view gray
  ready = [env.scalar_obs()[0, 0]]
  dead_steps = 1
  for _ in range(49):
    env.step(act())
    ready.append(env.scalar_obs()[0, 0])
    assert env.avatars()[1, 3] == 0 and env.scalar_obs()[1, 0] == 0.0     # readyToShoot is 0 when dead (:737-744)
    dead_steps += 1
  env.step(act())
  assert env.avatars()[1, 3] == 1 and dead_steps == 50                    # framesTillRespawn = 50 (clean_up.py:713)
  spawn_cells = {int(c) for c in tables[0]['spawn_cells_3']}
  x, y = env.avatars()[1][:2]
  assert y * W_CELLS + x in spawn_cells                                   # teleportToGroup('spawnPoints', ...)
  np.testing.assert_allclose(ready[:11], [0.0] + [0.1 * k for k in range(1, 11)], atol=1e-12)  # cooldownTime = 10


def test_zap_footprint_length3_radius1(oracle, clean_up_blob, tables):
  _, info = tables
  env = fresh(oracle, clean_up_blob)
  env.debug_set_avatar(0, 10, 12, N)
  env.step(act(p0=ZAP))
  cells = set(layer_sprites(env, info, 'beamZap'))
  want = {(10, 11), (10, 10), (10, 9), (9, 12), (9, 11), (9, 10), (11, 12), (11, 11), (11, 10)}
  assert cells == {y * W_CELLS + x for x, y in want}                     # SURVEY.md A.8 footprint
  env.step(act())
  assert not layer_sprites(env, info, 'beamZap')                          # hit sprites live one frame
  env.debug_set_avatar(0, 1, 12, W)                                      # facing the wall: nothing forward
  for _ in range(10):
    env.step(act())
  env.step(act(p0=ZAP))
  cells = set(layer_sprites(env, info, 'beamZap'))
  assert cells == {y * W_CELLS + x for x, y in {(1, 11), (1, 13)}}       # only the two lateral cells (walls block, no sprite)


def test_clean_beam_cleans_dirt_and_cumulant_lags_one_step(oracle, clean_up_blob, tables):
  _, info = tables
  env = fresh(oracle, clean_up_blob)
  env.debug_set_avatar(0, 3, 7, N)
  dirt_before = {c for c, s in layer_sprites(env, info, 'upperPhysical').items() if s == 'Dirt'}
  n0 = env.counters()['dirt']
  env.step(act(p0=CLEAN))
  dirt_after = {c for c, s in layer_sprites(env, info, 'upperPhysical').items() if s == 'Dirt'}
  cleaned = dirt_before - dirt_after
  n_events = sum(1 for e in env.events() if e[0] == 'player_cleaned')
  assert 1 <= len(cleaned) <= 3 and n_events == len(cleaned)              # one per column; dirt stops the ray
  assert env.counters()['dirt'] == n0 - len(cleaned)
  for c in cleaned:                                                       # each cleaned cell is the first dirt of its column
    x, y = c % W_CELLS, c // W_CELLS
    assert abs(x - 3) <= 1 and 7 - y <= 3 - abs(x - 3)
  assert (env.scalar_obs()[:, 1] == 0).all()                             # set during the drain of step t ...
  env.step(act())
  assert env.scalar_obs()[:, 1].tolist() == [0.0] + [1.0] * 6            # ... observed at t+1 by the OTHERS (A.3)
  env.step(act())
  assert (env.scalar_obs()[:, 1] == 0).all()


def test_apple_is_eaten_on_entry_same_frame(oracle, clean_up_blob, tables):
  sec, info = tables
  env = fresh(oracle, clean_up_blob)
  oid, cell, _ = (int(v) for v in sec['cu_apple'][40])
  x, y = cell % W_CELLS, cell // W_CELLS
  env.step(act())
  env.debug_set_object_state(oid, 0)                                     # 'apple' (clean_up.py:361-364)
  assert layer_sprites(env, info, 'upperPhysical').get(cell) == 'Apple'
  env.debug_set_avatar(0, x, y + 1 if y < 19 else y - 1, N if y < 19 else S)
  env.step(act(p0=FORWARD))
  assert env.rewards().tolist() == [1.0] + [0.0] * 6                     # Edible:onEnter -> Taste:consumed (role free)
  assert ('edible_consumed', 1, 0) in env.events()
  assert env.object_state(oid) == 1                                      # back to appleWait within the same update (A.4)
  assert cell not in layer_sprites(env, info, 'upperPhysical')
  env.step(act())
  assert (env.rewards() == 0).all()                                      # reward is per frame (Avatar:preUpdate)


def test_apple_growing_under_a_standing_avatar_is_eaten(oracle, clean_up_blob, tables):
  sec, _ = tables
  env = fresh(oracle, clean_up_blob)
  oid, cell, _ = (int(v) for v in sec['cu_apple'][40])
  env.debug_set_avatar(0, cell % W_CELLS, cell // W_CELLS, N)
  env.step(act())
  env.debug_set_object_state(oid, 0)                                     # contact is symmetric on placement (A.5)
  assert env.rewards()[0] == 1.0 and env.object_state(oid) == 1


def test_dirt_spawns_only_after_delay_at_rate_one_half(oracle, clean_up_blob):
  env = fresh(oracle, clean_up_blob, seed=11)
  for _ in range(50):
    env.step(act())
  assert env.counters()['dirt'] == 79                                    # delayStartOfDirtSpawning = 50
  for _ in range(60):
    env.step(act())
  grown = env.counters()['dirt'] - 79
  assert 15 <= grown <= 45                                               # Binomial(60, 0.5)
  for _ in range(400):
    env.step(act())
  assert env.counters() ['dirt'] == 147 and env.counters()['clean'] == 0  # river saturates


def test_apple_growth_follows_dirt_fraction(oracle, clean_up_blob, tables):
  sec, info = tables
  env = fresh(oracle, clean_up_blob, seed=5)
  for _ in range(40):
    env.step(act())
  assert not [s for s in layer_sprites(env, info, 'upperPhysical').values() if s == 'Apple']  # 79/147 > 0.4 -> p = 0
  for oid, cell, dirty in sec['cu_dirt']:
    env.debug_set_object_state(int(oid), 0)                              # everything clean: dirtWait
  assert env.counters()['dirt'] == 0
  env.step(act())
  apples = [s for s in layer_sprites(env, info, 'upperPhysical').values() if s == 'Apple']
  assert 1 <= len(apples) <= 18                                          # Binomial(122, 0.05), mean 6.1


def test_episode_ends_on_interval_boundaries_and_auto_resets(oracle, clean_up_blob):
  lengths = []
  for seed in range(12):
    env = oracle.OracleEnv(clean_up_blob, 100 + seed)
    env.reset()
    n = 0
    while True:
      n += 1
      st = env.step(act())
      if st == 2:
        break
      assert st == 1 and env.discount() == 1.0
    assert env.discount() == 0.0                                         # dm_env.termination
    assert n >= 1099 and (n + 1) % 100 == 0 and n <= 5000                # minimumFramesPerEpisode 1000, interval 100
    lengths.append(n)
    assert env.step(act()) == 0 and env.counters()['episode'] == 1        # A.17: step after LAST -> FIRST
    assert (env.rewards() == 0).all() and env.discount() == 0.0 and env.counters()['step'] == 0
  assert len(set(lengths)) > 1                                            # p = 0.2 per interval


def test_render_geometry_and_sprite_maps(oracle, clean_up_blob, tables):
  _, info = tables
  env = fresh(oracle, clean_up_blob)
  env.debug_set_avatar(0, 3, 10, W)
  env.debug_set_avatar(1, 3, 12, N)
  env.step(act())
  rgb = env.rgb()
  # 9 cells forward of x=3 facing W leaves the map after 3 cells: the top 6 cell rows are OutOfBounds black.
  assert (rgb[0][:6 * 8] == 0).all() and (rgb[0][6 * 8:7 * 8] != 0).any()
  # clean_up.py:487-493,664-665: 'Self' takes human_readable[0] (blue); avatar i gets the (i+1)-th colour.
  blue, purple, pink = (45, 110, 220), (125, 50, 200), (205, 5, 165)
  own = rgb[0][9 * 8:10 * 8, 5 * 8:6 * 8].reshape(-1, 3)                  # viewer sits at column left=5, row forward=9
  assert (own == blue).all(1).any() and not (own == purple).all(1).any()  # spriteMap Avatar1 -> Self
  world = env.world_rgb()
  cell0 = world[10 * 8:11 * 8, 3 * 8:4 * 8].reshape(-1, 3)
  assert (cell0 == purple).all(1).any() and not (cell0 == blue).all(1).any()  # WORLD.RGB shows true colours
  cell1 = world[12 * 8:13 * 8, 3 * 8:4 * 8].reshape(-1, 3)
  assert (cell1 == pink).all(1).any()
  own1 = rgb[1][9 * 8:10 * 8, 5 * 8:6 * 8].reshape(-1, 3)
  assert (own1 == blue).all(1).any() and not (own1 == pink).all(1).any()  # spriteMap Avatar2 -> Self
  # avatar 0 stands two cells in front of avatar 1 (facing N): seen there in its true colour.
  other = rgb[1][7 * 8:8 * 8, 5 * 8:6 * 8].reshape(-1, 3)
  assert (other == purple).all(1).any()


def test_water_animation_flips_every_two_frames(oracle, clean_up_blob, tables):
  sec, info = tables
  env = fresh(oracle, clean_up_blob)
  cells = [int(c) for c in sec['cu_water'][:, 1]]
  def phases():
    g = env.grid()[info['layers'].index('background')]
    return [info['sprites'][(int(g[c]) - 1) // 4] for c in cells]
  p = []
  for _ in range(5):
    env.step(act())
    p.append(phases())
  assert p[0] != p[1] and p[1] == p[2] and p[2] != p[3] and p[3] == p[4]  # switches at frames 2, 4 (gameFramesPerAnimationFrame=2)
  assert len(set(p[0])) > 1                                                # randomStartFrame


def test_determinism_and_seed_sensitivity(oracle, clean_up_blob):
  def trace(seed):
    env = oracle.OracleEnv(clean_up_blob, seed)
    env.reset()
    rng = np.random.default_rng(0)
    out = []
    for _ in range(60):
      env.step(rng.integers(0, 9, P))
      out.append((env.avatars().tobytes(), env.world_rgb().tobytes()))
    return out
  assert trace(7) == trace(7)                                              # builder_test.py:47-70
  assert trace(7) != trace(8)
