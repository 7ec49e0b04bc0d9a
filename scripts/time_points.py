#!/usr/bin/env python
"""Times the materialising point-cloud kernels (dm_points.cu) on config-2 sized inputs: GB/s of tensor bytes."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dungeon_maps_b200 as dmap
from dungeon_maps_b200 import synth
dev = torch.device("cuda", 0)
b, H, W = 16, 480, 640
depth = synth.iid_depth(b, H, W, device=dev)
fx = fy = 320 / math.tan(math.radians(35)); cx, cy = 320., 240.
def timeit(f, nbytes, name, n=20):
  for _ in range(3): f()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): f()
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  print(f"{name}: {ms*1e3:.1f} us, {nbytes/ms/1e6:.0f} GB/s of tensor bytes")
N = b * H * W
pc, valid = dmap.depth_map_to_point_cloud(depth, None, fx, fy, cx, cy, 0.15, 5.05)
timeit(lambda: dmap.depth_map_to_point_cloud(depth, None, fx, fy, cx, cy, 0.15, 5.05), N * (4 + 12 + 1), "depth_map_to_point_cloud")
pose = synth.poses(b, 1)
timeit(lambda: dmap.camera_to_local_space(pc, cam_pitch=-0.17, cam_height=0.88), N * 24, "camera_to_local_space")
timeit(lambda: dmap.local_to_global_space(pc, cam_pose=pose), N * 24, "local_to_global_space")
timeit(lambda: dmap.camera_to_image_space(pc, focal_x=fx, focal_y=fy, center_x=cx, center_y=cy, height=H), N * 24, "camera_to_image_space")
