"""The renderer's lane -> cell dealing (make_lane_map in csrc/engine.cu): a permutation of the strip's (row, cell) items
whose 64-bit shared-memory stores are bank-conflict free per half-warp. Pure host code: runs without a GPU."""

import ctypes

import pytest

from meltingpot_b200 import engine


def _lane_map(n_rows, n_cells, pitch_slots, iters, scattered=1):
  out = (ctypes.c_uint32 * 32)()
  rc = engine.load_library().mp_debug_lane_map(n_rows, n_cells, pitch_slots, iters, scattered, out)
  return rc, list(out)


def _extra_wavefronts(table, n_rows, n_cells, pitch, iters):
  extra = 0
  for it in range(iters):
    for half in range(2):
      cnt = {}
      for lane in range(16 * half, 16 * half + 16):
        cell = (table[lane] >> (6 * it)) & 63
        if cell != 63:
          pair = (pitch * (lane % n_rows) + 3 * cell) % 16
          cnt[pair] = cnt.get(pair, 0) + 1
      extra += max(cnt.values(), default=1) - 1
  return extra


@pytest.mark.parametrize('n_rows,n_cells', [(8, 11)] + [(r, w) for w in (21, 24, 27, 30, 39) for r in (4, 2)])
def test_default_dealing_keeps_cells_whole_and_beats_the_plain_order(n_rows, n_cells):
  G = 32 // n_rows
  iters = -(-n_cells // G)
  pitch = 3 * n_cells
  rc, table = _lane_map(n_rows, n_cells, pitch, iters, scattered=0)
  assert rc == 0
  cells = []
  for it in range(iters):
    for g in range(G):
      group = {(table[g * n_rows + j] >> (6 * it)) & 63 for j in range(n_rows)}
      assert len(group) == 1  # all pixel rows of a cell on one lane group, in one turn
      cells += [c for c in group if c != 63]
  assert sorted(cells) == list(range(n_cells))
  plain = [sum(min(63, l // n_rows + G * i) << (6 * i) for i in range(iters)) for l in range(32)]
  plain = [sum((((p >> (6 * i)) & 63) if ((p >> (6 * i)) & 63) < n_cells else 63) << (6 * i) for i in range(iters)) for p in plain]
  assert _extra_wavefronts(table, n_rows, n_cells, pitch, iters) <= _extra_wavefronts(plain, n_rows, n_cells, pitch, iters)
  if (n_rows, n_cells) == (8, 11):
    assert _extra_wavefronts(plain, 8, 11, pitch, iters) == 5 and _extra_wavefronts(table, 8, 11, pitch, iters) == 2


@pytest.mark.parametrize('n_rows,n_cells', [(8, 11), (8, 9), (8, 13), (8, 16)] + [(r, w) for w in (16, 18, 21, 23, 24, 25, 27, 30, 39, 40) for r in (4, 2)])
def test_dealing_is_a_conflict_free_permutation(n_rows, n_cells):
  per_turn = 32 // n_rows
  iters = -(-n_cells // per_turn)
  pitch = 3 * n_cells  # 24-byte cells, 8-byte slots
  rc, table = _lane_map(n_rows, n_cells, pitch, iters)
  assert rc == 0
  seen = set()
  for it in range(iters):
    for half in range(2):
      banks = set()
      for lane in range(16 * half, 16 * half + 16):
        cell = (table[lane] >> (6 * it)) & 63
        if cell == 63:
          continue
        assert cell < n_cells
        row = lane % n_rows
        assert (row, cell) not in seen
        seen.add((row, cell))
        pair = (pitch * row + 3 * cell) % 16  # first 8-byte slot of the lane's 24 bytes, as a bank pair
        assert pair not in banks, (it, half, lane)
        banks.add(pair)
  assert len(seen) == n_rows * n_cells  # every pixel row of every cell is drawn exactly once
  # 128-bit atlas loads are served per quarter-warp and hit bank group (pixel row mod 8): with 8-row strips the eight
  # lanes of a quarter-warp hold eight different rows by construction
  if n_rows == 8:
    for q in range(4):
      assert sorted(l % 8 for l in range(8 * q, 8 * q + 8)) == list(range(8))
