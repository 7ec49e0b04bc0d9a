#!/usr/bin/env python
"""A/B of the builder workloads of bench.py under a tile-layout knob (dm_debug_set_tile_rows):
usage: python scripts/time_builder.py builder_fixed -1 -3"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dungeon_maps_b200 import _native as nat
key = sys.argv[1]
if os.environ.get("DM_DENSE") is not None:
  nat.lib().dm_debug_set_dense_shift(int(os.environ["DM_DENSE"]))
dev = torch.device("cuda", 0)
args = types.SimpleNamespace(scene="room")
frames = None
for rows in [int(x) for x in sys.argv[2:]]:
  nat.lib().dm_debug_set_tile_rows(rows)
  w = bench.BuilderWorkload(args, key)
  w.t = 0
  w.setup(dev, 0, frames)
  frames = w.frames
  w.reset_counters()
  for _ in range(w.EPISODE): w.step()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(w.EPISODE): w.step()
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / w.EPISODE
  print(key, "rows", rows, "ms/step %.4f" % ms, "env-steps/s %.0f" % (w.B / ms * 1e3))
nat.lib().dm_debug_set_tile_rows(-1)
