"""Which cell stacks take the renderer's multi-sprite path, and how often (diagnostic; needs a GPU)."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from meltingpot_b200 import blob as mb, engine, substrates

name, players, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
blob = substrates.load_blob(name, ('default',) * players)
info = json.loads(mb.section_text(mb.unpack(blob), 'info_json'))
names = info['sprites']
eng = engine.Engine(blob, B, seed=1)
eng.reset()
gen = torch.Generator(device='cuda').manual_seed(0)
for t in range(steps):
  eng.step_state(torch.randint(0, eng.num_actions, (B, players), generator=gen, device='cuda', dtype=torch.int32))
pair, flags = eng.render_tables()
pair = pair.astype(np.int64); flags = flags.astype(np.int64)
n_total = len(flags)
grid = eng.grid.cpu().numpy().astype(np.int64) & 0xffff  # [B, L, cells_pad]
L = grid.shape[1]
cells = eng.buffers.grid_cells if hasattr(eng, 'buffers') else grid.shape[2]
cnt = collections.Counter()
n_cells = 0
nm = lambda s: names[s] if s < len(names) else 'merged%d' % s
for b in range(min(B, 16)):
  for c in range(cells):
    col = grid[b, :, c]
    n_cells += 1
    lo = 0
    for l in range(L - 1, 0, -1):
      if col[l] and (flags[(col[l] - 1) >> 2] & 1): lo = l; break
    rec, cur, merging = [], 0, True
    for l in range(lo, L):
      v = col[l]
      if not v: continue
      if cur == 0: cur = v; continue
      if merging:
        m = pair[(cur - 1) >> 2, (v - 1) >> 2]
        if m and ((cur - 1) ^ (v - 1)) & 3 == 0: cur = 1 + m * 4 + ((v - 1) & 3); continue
        merging = False
      rec.append(cur); cur = v
    if cur: rec.append(cur)
    if len(rec) == 1 and flags[(rec[0] - 1) >> 2] & 1: continue
    cnt[tuple(nm((v - 1) >> 2) for v in rec)] += 1
slow = sum(cnt.values())
print(json.dumps({'substrate': name, 'n_total': int(n_total), 'slow_fraction': slow / n_cells}))
for k, v in cnt.most_common(40): print('%6.3f%%  %s' % (100.0 * v / n_cells, ' + '.join(k)))
