#!/usr/bin/env python
"""Experiment helper (needs a GPU): the merge of the config-4 walk's last step, timed with CUDA events, with and
without the per-plane rectangles of the world map (DmFuseSource.plane_box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dungeon_maps_b200 as dmap


class A: pass
args = A(); args.scene = "room"
wl = bench.BuilderWorkload(args, "builder")
wl.setup(torch.device("cuda", 0), 0)
for _ in range(99):
  wl.step()
wm = wl.builder.world_map
tb = wm._tracked_box
print("world", tuple(wm.mask.shape), "plane_box", None if tb.plane_box is None else tb.plane_box[:4].tolist())
local = wl.builder.plot(wl.frames[99], cam_pose=wl.poses[99], **wl.local_kw)
target = wl.builder.proj.clone(cam_pose=wl.poses[99])


def timed(label):
  for _ in range(3):
    out = dmap.fuse_topdown_maps(wm, local, map_projector=target)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20):
    out = dmap.fuse_topdown_maps(wm, local, map_projector=target)
  e1.record()
  torch.cuda.synchronize()
  print(f"{label}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per merge (host sync included)")
  return out


a = timed("with rectangles")
planes, tb.plane_box = tb.plane_box, None
b = timed("full planes    ")
tb.plane_box = planes
assert torch.equal(a.topdown_map, b.topdown_map) and torch.equal(a.mask, b.mask)
