#!/usr/bin/env python
"""Experiment helper (not part of the product): times the HOST-buffer entry (dm_orth_project_host_f32) on the
config-2 workload for several pipeline chunk sizes ($DM_HOST_CHUNK), next to the raw pinned-copy rates of the box
(H2D alone, D2H alone, both directions at once) — the floor of the e2e number.
usage: python scripts/time_e2e.py [--chunks 1,2,4,8] [--steps 5]"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dungeon_maps_b200 import hostapi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", default="0,1,2,4,8")
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
B, H, W, C, MH, MW = 64, 480, 640, 16, 400, 400
dev = torch.device("cuda", 0)
depth, values, pose = synth.frames("room", B, H, W, C, seed=0, device=dev)
h_depth, h_values = depth.cpu().pin_memory().numpy(), values.cpu().pin_memory().numpy()
del depth, values
out = (torch.empty((B, C, MH, MW), dtype=torch.float32).pin_memory().numpy(),
       torch.empty((B, C, MH, MW), dtype=torch.uint8).pin_memory().numpy(),
       torch.empty((B, 1, MH, MW), dtype=torch.float32).pin_memory().numpy())
intr_f = (W / 2.) / math.tan(math.radians(70) / 2.)
kw = dict(map_res=0.03, map_width=MW, map_height=MH, focal_x=intr_f, focal_y=intr_f, center_x=(W - 1) / 2.,
          center_y=(H - 1) / 2., trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None, clip_border=10,
          to_global=False, fill_value=-np.inf, get_height_map=True)
h2d = h_depth.nbytes + h_values.nbytes
d2h = sum(o.nbytes for o in out)

# raw copy rates
src = torch.from_numpy(h_values)
dst = torch.empty_like(src, device=dev)
back = torch.from_numpy(out[0])
dsrc = torch.empty_like(back, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def rate(fn, nbytes, reps=3):
  fn(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(reps):
    fn()
  torch.cuda.synchronize()
  return nbytes * reps / (time.perf_counter() - t0) / 1e9
def both():
  with torch.cuda.stream(s1):
    dst.copy_(src, non_blocking=True)
  with torch.cuda.stream(s2):
    back.copy_(dsrc, non_blocking=True)
res = {"h2d_alone_GBs": rate(lambda: dst.copy_(src, non_blocking=True), src.nbytes),
       "d2h_alone_GBs": rate(lambda: back.copy_(dsrc, non_blocking=True), back.nbytes)}
t = rate(both, 1)  # seconds-based below
both(); torch.cuda.synchronize()
t0 = time.perf_counter(); both(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
res["duplex_ms_for_%.2fGB_in_%.2fGB_out" % (src.nbytes / 1e9, back.nbytes / 1e9)] = dt * 1e3
res["floor_ms_per_step"] = max(h2d / res["h2d_alone_GBs"], d2h / res["d2h_alone_GBs"]) / 1e6
print(json.dumps(res))
del dst, dsrc

for ch in a.chunks.split(","):
  if ch == "0":
    os.environ.pop("DM_HOST_CHUNK", None)
  else:
    os.environ["DM_HOST_CHUNK"] = ch
  step = lambda: hostapi.orth_project_host(h_depth, h_values, None, pose.cpu(), 200., 0., math.radians(-10), 0.88,
                                           device=0, out=out, **kw)
  step(); step()
  ts = []
  for _ in range(a.steps):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
  print(json.dumps({"chunk": ch, "ms_best": min(ts) * 1e3, "ms_mean": sum(ts) / len(ts) * 1e3,
                    "maps_per_s": B / (sum(ts) / len(ts)), "h2d_GBs": h2d / min(ts) / 1e9, "d2h_GBs": d2h / min(ts) / 1e9}))
