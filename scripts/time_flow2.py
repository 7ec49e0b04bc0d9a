#!/usr/bin/env python
"""camera_affine_grid: does the step time depend on where the 629 MB output lands?  (same block every call / two blocks
alternating / fresh pose deltas)"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dungeon_maps_b200 as dmap
from dungeon_maps_b200 import synth
b, H, W = 256, 480, 640
dev = torch.device("cuda", 0)
depth, _, pose = synth.frames("room", b, H, W, 0, seed=0, device=dev)
proj = dmap.MapProjector(width=W, height=H, hfov=math.radians(70), cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                         cam_pitch=math.radians(-10), cam_height=0.88, map_res=0.03, map_width=400, map_height=400, device=dev)
delta = (pose * 0.1).cpu()
deltas = [(synth.poses(b, i) * 0.1).cpu() for i in range(64)]
def run(name, hold, fresh):
  out = None
  for i in range(5): out = proj.camera_affine_grid(depth, deltas[i] if fresh else delta)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(100):
    if hold: out = proj.camera_affine_grid(depth, deltas[i % 64] if fresh else delta)
    else: proj.camera_affine_grid(depth, deltas[i % 64] if fresh else delta)
  e1.record(); torch.cuda.synchronize()
  print(name, "ms/step %.4f" % (e0.elapsed_time(e1) / 100))
run("same block, same poses ", False, False)
run("two blocks, same poses ", True, False)
run("same block, fresh poses", False, True)
run("two blocks, fresh poses", True, True)
