"""Prints details of the first GPU-vs-oracle pixel mismatch (run on a GPU box)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from meltingpot_b200 import engine, substrates, blob as blob_lib
from oracle import binding as ob

blob = substrates.load_blob('clean_up')
info = json.loads(blob_lib.section_text(blob_lib.unpack(blob), 'info_json'))
B, seed, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1, 300
flags = 3 | (int(os.environ.get('MP_DEBUG_CELL', '0')) << 8)
eng = engine.Engine(blob, B, seed=seed, flags=flags)
envs = [ob.OracleEnv(blob, seed + b) for b in range(B)]
rng = np.random.default_rng(0)
eng.reset(); [e.reset() for e in envs]
W = 30
for t in range(steps):
  acts = rng.integers(0, 9, size=(B, 7)).astype(np.int32)
  eng.step(torch.from_numpy(acts).cuda()); torch.cuda.synchronize()
  rgb = eng.rgb.cpu().numpy(); world = eng.world_rgb.cpu().numpy()
  grid = eng.grid.cpu().numpy().view(np.uint16)
  found = False
  for b, e in enumerate(envs):
    e.step(acts[b])
    o_rgb, o_world = e.rgb(), e.world_rgb()
    if not np.array_equal(o_world, world[b]):
      d = np.argwhere((o_world != world[b]).any(-1))
      cells = sorted({(int(y) // 8, int(x) // 8) for y, x in d})
      print(f'WORLD mismatch step {t} env {b}: {len(d)} px, cells(y,x) {cells[:20]}')
      for (cy, cx) in cells[:3]:
        c = cy * W + cx
        print('  cell', (cx, cy), 'stack', [(info['layers'][l], info['sprites'][(int(v) - 1) // 4], (int(v) - 1) % 4) for l, v in enumerate(grid[b][:, c]) if v])
        print('  oracle row0', o_world[cy * 8, cx * 8:cx * 8 + 8].tolist())
        print('  gpu    row0', world[b][cy * 8, cx * 8:cx * 8 + 8].tolist())
      found = True
    for p in range(7):
      if not np.array_equal(o_rgb[p], rgb[b, p]):
        d = np.argwhere((o_rgb[p] != rgb[b, p]).any(-1))
        cells = sorted({(int(y) // 8, int(x) // 8) for y, x in d})
        av = e.avatars()[p]
        print(f'RGB mismatch step {t} env {b} player {p} avatar {av.tolist()}: {len(d)} px, view cells(vy,vx) {cells[:20]}')
        for (vy, vx) in cells[:3]:
          print('  oracle row0', o_rgb[p][vy * 8, vx * 8:vx * 8 + 8].tolist())
          print('  gpu    row0', rgb[b, p][vy * 8, vx * 8:vx * 8 + 8].tolist())
        found = True
  if found:
    break
else:
  print('no mismatch in', steps, 'steps')
