#!/usr/bin/env python
"""Experiment helper (not part of the product): times dm_orth_project_f32 on the bench workload with the
library named by $DM_B200_LIB (a variant built by scripts/exp_build.sh), and prints the DM_PROFILE cycle
counters when the variant carries them.
usage: DM_B200_LIB=build/exp/lib_x.so python scripts/time_proj.py [--scene room|iid] [--steps 50] [--fill ninf|0]"""
import argparse, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dungeon_maps_b200 as dmap
from dungeon_maps_b200 import maps as dmaps, synth

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="room")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--fill", default="ninf")
ap.add_argument("--b", type=int, default=64)
ap.add_argument("--c", type=int, default=16)
ap.add_argument("--hw", default="480x640")
ap.add_argument("--rows", type=int, default=-1, help="tile layout (dm_debug_set_tile_rows): -1 auto, 0 row tiles, 4 / 8 2-D tiles")
a = ap.parse_args()
H, W = map(int, a.hw.split("x"))
dev = torch.device("cuda", 0)
from dungeon_maps_b200 import _native as _nat
_nat.lib().dm_debug_set_tile_rows(a.rows)
depth, values, pose = synth.frames(a.scene, a.b, H, W, a.c, seed=0, device=dev)
proj = dmap.MapProjector(width=W, height=H, hfov=math.radians(70), cam_pose=[0., 0., 0.], width_offset=200.,
                         height_offset=0., cam_pitch=math.radians(-10), cam_height=0.88, map_res=0.03,
                         map_width=400, map_height=400, trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10,
                         to_global=False, fill_value=dmap.NINF if a.fill == "ninf" else 0.0, device=dev)
pose_host = pose.cpu()
step = lambda: proj.orth_project(depth, values if a.c else None, cam_pose=pose_host, get_height_map=True)
for _ in range(5):
  step()
torch.cuda.synchronize()
ws = [w for w in dmaps._workspaces.values()]
prof = os.environ.get("DM_PROFILE") == "1"
if prof:
  for w in ws:
    w.view(torch.int32)[320:384].zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
  step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
out = {"lib": os.environ.get("DM_B200_LIB", "default"), "rows": a.rows, "hw": a.hw, "c": a.c, "scene": a.scene, "fill": a.fill, "ms_per_step": round(ms, 4),
       "maps_per_s": round(a.b / ms * 1e3)}
if prof:
  c = ws[0].view(torch.int64)[160:184].cpu().tolist()
  n = a.steps
  if os.environ.get("DM_PROJ_KERNEL", "rl")[0] == "w":
    names = ["A", "B1", "B2", "runlets", "waitfull_proj", "waitfull_res", "resolve", "n_proj",
             "prod_claim", "prod_wait_empty", "prod_block", "prod_issue"]
    d = dict(zip(names, c))
    npj = max(d["n_proj"], 1)
    out["per_proj_tile_cycles(warp0)"] = {k: round(d[k] / npj) for k in ("A", "B1", "B2", "waitfull_proj")}
    out["runlets_per_warp_tile"] = round(d["runlets"] / npj, 1)
    out["raw_per_step"] = {k: round(v / n) for k, v in d.items()}
  else:
    names = ["A", "B1", "publish", "next", "depwait", "B2", "n", "_7", "depwait", "load", "publish", "next", "store", "n", "occupied"]
    npj, nrs = max(c[6], 1), max(c[13], 1)
    out["proj_ticket_cycles(warp0)"] = {names[i]: round(c[i] / npj) for i in range(6)}
    out["proj_ticket_total"] = round(sum(c[:6]) / npj)
    out["resolve_ticket_cycles(warp0)"] = {names[i]: round(c[i] / nrs) for i in range(8, 13)}
    out["resolve_ticket_total"] = round(sum(c[8:13]) / nrs)
    out["tickets_per_warp0_per_step"] = {"proj": c[6] / n / 592, "resolve": c[13] / n / 592, "occupied_frac": c[14] / nrs}
print(json.dumps(out))
if os.environ.get("DM_SHOW_HINT"):
  print("density hint ctrl[4] =", int(ws[0].view(torch.int32)[4]), "-> touched fraction", (int(ws[0].view(torch.int32)[4]) - 1) / 65536)
