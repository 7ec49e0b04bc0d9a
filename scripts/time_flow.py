#!/usr/bin/env python
"""Experiment helper (not part of the product): times dm_affine_grid_f32 on the config-3 workload with the library
named by $DM_B200_LIB (scripts/exp_build.sh).  usage: DM_B200_LIB=build/exp/lib_x.so python scripts/time_flow.py"""
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
for _ in range(5):
  proj.camera_affine_grid(depth, delta)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
  proj.camera_affine_grid(depth, delta)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 100
print(json.dumps({"lib": os.environ.get("DM_B200_LIB", "default"), "ms_per_step": round(ms, 4), "GBps": round(b * H * W * 12 / ms / 1e6)}))
