"""The renderer's lane -> cell dealing (make_lane_map in csrc/engine.cu): a permutation of the strip's (row, cell) items
whose 64-bit shared-memory stores are bank-conflict free per half-warp. Pure host code: runs without a GPU."""

import ctypes

import pytest

from meltingpot_b200 import engine


def _lane_map(n_rows, n_cells, pitch_slots, iters):
  out = (ctypes.c_uint32 * 32)()
  rc = engine.load_library().mp_debug_lane_map(n_rows, n_cells, pitch_slots, iters, out)
  return rc, list(out)


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
