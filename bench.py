#!/usr/bin/env python
"""bench.py — top-down maps/sec of the fused projection (BASELINE.json config 2), plus the other
workloads of the path as separately labelled lines.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload proj|flow|builder|builder_fixed|proj5] [--scene room|iid]

Default (`--workload proj`, the line the driver reads): one "step" = one pass of the hot path
(dm_orth_project_f32: fused projection + resolve) over one batch of 64 synthetic 480x640 depth frames
+ 16 one-hot semantic channels → 400x400 local maps (topdown, mask, height).
Other workloads (BASELINE.json configs 3, 4, 5; results kept under profiles/):
  flow           camera_affine_grid, 256 x 480x640 frames with random pose deltas per step
  builder        MapBuilder.step (plot + reference-parity merge), 32 environments walking for 100 steps
  builder_fixed  the same walk merged in place into fixed 2400x2400 world canvases (opt-in mode)
  proj5          the projection at 1280x720 with 40 semantic channels, one 64-frame chunk per step
N > 1: one process per GPU (torchrun; `python bench.py --gpus N` on its own re-launches itself that way), every
rank works on its own environments, no collective on the data path (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import numpy as np
import torch

HFOV, PITCH, CAM_H, RES = math.radians(70), math.radians(-10), 0.88, 0.03
MH, MW = 400, 400


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
  FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
            "clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index: int):
    self.rows, self.proc = [], None
    try:
      self.proc = subprocess.Popen(
        ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(gpu_index)],
        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._pump, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.rows.append((time.perf_counter(), line.strip()))

  def wait_ready(self, timeout: float = 3.0) -> None:
    """nvidia-smi needs a moment before its first sample; short timed regions would otherwise see none."""
    t_end = time.perf_counter() + timeout
    while self.proc is not None and not self.rows and time.perf_counter() < t_end:
      time.sleep(0.01)

  def stop(self, t0: float, t1: float) -> dict:
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    inside = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows[-3:]]
    sm, smax, reasons = [], [], set()
    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    for r in inside:
      f = [x.strip() for x in r.split(",")]
      try:
        sm.append(float(f[0])); smax.append(float(f[1]))
      except Exception:
        continue
      for n, v in zip(names, f[3:7]):
        if v.lower().startswith("active"):
          reasons.add(n)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  try:
    with open(path) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  except Exception:
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key: str):
  """dram bytes per launch of the workload's dominant kernel from the committed ncu --set full summary."""
  try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
      d = json.load(f)
    if key == "proj":
      return d.get("dram_bytes_per_step")
    return (d.get(key) or {}).get("dram_bytes_per_step")
  except Exception:
    return None


def intrinsics(W, H):
  cx, cy = W / 2., H / 2.
  fx = cx / np.tan(HFOV / 2.)
  return fx, fx, cx, cy


# ================================================================================================
# Workloads.  Each one: setup(dev, rank) once; step() = one pass over one batch on the device;
# e2e_setup()/e2e_step() = the same pass from pinned HOST buffers with the copies inside;
# cpu_step(threads) -> units processed by one bounded CPU pass of the oracle port.
# ================================================================================================

class ProjWorkload:
  """BASELINE config 2 (and config 5's shapes with key='proj5')."""
  unit = "maps/s"
  kernel = "dm::proj_ws_kernel (one step = ONE persistent launch of dm_orth_project_f32: projection + resolve)"

  def __init__(self, args, key="proj"):
    self.key, self.scene = key, args.scene
    self.labels = key.endswith("_labels")
    key = key.replace("_labels", "")
    if key == "proj":
      self.B, self.H, self.W, self.C = 64, 480, 640, 16
      self.metric = "top-down maps/sec (batch 64, 640x480 depth + 16 semantic channels -> 400x400 maps)"
      self.name = ("BASELINE config 2: 64 x 480x640 depth + 16-channel one-hot semantics -> 400x400 maps "
                   "(topdown + mask + height), per GPU")
    else:
      self.B, self.H, self.W, self.C = 64, 720, 1280, 40
      self.metric = "top-down maps/sec (1280x720 depth + 40 semantic channels -> 400x400 maps, 64-frame chunks)"
      self.name = ("BASELINE config 5 (steady state of the 4096-frame job): one 64-frame chunk of 1280x720 depth + "
                   "40-channel one-hot semantics -> 400x400 maps per step, per GPU; every GPU streams 4096/N frames")
    self.units_per_step = self.B
    # read 4*N*(1+C) input bytes, write Mh*Mw*(4C + C + 4) output bytes per frame
    in_bytes = self.H * self.W * (5 if self.labels else 4 * (1 + self.C))
    self.algo_bytes_per_step = self.B * (in_bytes + MH * MW * (4 * self.C + self.C + 4))
    self.l2_note = (f"{self.algo_bytes_per_step / 1e9:.2f} GB move per step (inputs {self.B * in_bytes / 1e9:.2f} GB), larger than "
                    "the 126 MB L2; no flush needed")
    if self.labels:
      self.kernel = "dm::proj_lbl_kernel (one step = ONE persistent launch of dm_orth_project_labels_f32)"
      self.metric += " [semantics given as uint8 class ids]"
      self.name += "; semantics enter as (b,1,H,W) uint8 class ids (label_map=, num_classes=16): bit-identical outputs"

  def kwargs(self):
    fx, fy, cx, cy = intrinsics(self.W, self.H)
    return dict(map_res=RES, map_width=MW, map_height=MH, focal_x=fx, focal_y=fy, center_x=cx, center_y=cy,
                trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None, clip_border=10,
                to_global=False, flip_h=True, fill_value=-np.inf, reduction=None, get_height_map=True)

  def config(self, world):
    return {"workload": self.name, "scene": self.scene, "fill_value": "-inf", "frames_per_step_per_gpu": self.B,
            "parallelism": f"batch-sharded x{world}, no collective", "l2": self.l2_note}

  def setup(self, dev, rank):
    import dungeon_maps_b200 as dmap
    from dungeon_maps_b200 import synth
    self.dev = dev
    self.depth, self.values, pose = synth.frames(self.scene, self.B, self.H, self.W, self.C, seed=rank, device=dev)
    self.proj = dmap.MapProjector(width=self.W, height=self.H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=200.,
                                  height_offset=0., cam_pitch=PITCH, cam_height=CAM_H, map_res=RES, map_width=MW,
                                  map_height=MH, trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10,
                                  to_global=False, fill_value=dmap.NINF, device=dev)
    self.pose_host = pose.cpu()
    self.label_ids = self.values.argmax(1, keepdim=True).to(torch.uint8)

  def step(self):
    if self.labels:
      self.out = self.proj.orth_project(self.depth, cam_pose=self.pose_host, get_height_map=True,
                                        label_map=self.label_ids, num_classes=self.C)
    else:
      self.out = self.proj.orth_project(self.depth, self.values, cam_pose=self.pose_host, get_height_map=True)
    return self.out

  def e2e_setup(self):
    B, C = self.B, self.C
    self.h_depth = self.depth.cpu().pin_memory().numpy()
    if self.labels:
      self.h_values = self.label_ids.cpu().pin_memory().numpy()
    else:
      self.h_values = self.values.cpu().pin_memory().numpy()
    self.o_top = torch.empty((B, C, MH, MW), dtype=torch.float32).pin_memory().numpy()
    self.o_mask = torch.empty((B, C, MH, MW), dtype=torch.uint8).pin_memory().numpy()
    self.o_hgt = torch.empty((B, 1, MH, MW), dtype=torch.float32).pin_memory().numpy()
    self.h2d = self.h_depth.nbytes + self.h_values.nbytes + B * 192
    self.d2h = self.o_top.nbytes + self.o_mask.nbytes + self.o_hgt.nbytes
    self.e2e_path = ("hostapi.orth_project_host -> dm_orth_project_labels_host_f32 (pinned host buffers)" if self.labels
                     else "hostapi.orth_project_host -> dm_orth_project_host_f32 (pinned host buffers)")

  def e2e_step(self):
    from dungeon_maps_b200 import hostapi
    if self.labels:
      hostapi.orth_project_host(self.h_depth, None, None, self.pose_host, 200., 0., PITCH, CAM_H,
                                device=self.dev.index, out=(self.o_top, self.o_mask, self.o_hgt),
                                label_map=self.h_values, num_classes=self.C, **self.kwargs())
    else:
      hostapi.orth_project_host(self.h_depth, self.h_values, None, self.pose_host, 200., 0., PITCH, CAM_H,
                                device=self.dev.index, out=(self.o_top, self.o_mask, self.o_hgt), **self.kwargs())

  def e2e_check(self):
    assert np.array_equal(self.o_top, self.out[0].cpu().numpy()) and \
        np.array_equal(self.o_hgt, self.out[2][:, :1].cpu().numpy()), "host-buffer path and device path disagree"

  def cpu_setup(self, threads):
    from dungeon_maps_b200 import synth
    from oracle import dm_oracle as orc
    n = min(self.B, max(threads, 8))  # bounded sample: one frame per host thread, at most the batch
    depth, values, pose = synth.frames(self.scene, n, self.H, self.W, self.C, seed=0)
    d, v, p = depth.numpy(), values.numpy(), pose.numpy()
    kw = self.kwargs()
    orc.orth_project(d[:1], v[:1], None, p[:1], 200., 0., PITCH, CAM_H, threads=1, **kw)  # page-in
    self.cpu_units = n
    self.cpu_fn = lambda: orc.orth_project(d, v, None, p, 200., 0., PITCH, CAM_H, threads=threads, **kw)
    self.cpu_sample = (f"{n} frames per pass of the same workload ({self.scene} scene), oracle/dm_oracle.c "
                       f"(scalar C restatement of the reference CPU path) with {threads} OpenMP threads")


class FlowWorkload:
  """BASELINE config 3: compute_ego_flow / camera_affine_grid, batch 256 x 480x640, random pose deltas."""
  unit = "frames/s"
  key = "flow"
  kernel = "dm::flow_kernel (one step = one launch of dm_affine_grid_f32)"
  metric = "ego-flow frames/sec (camera_affine_grid, batch 256, 640x480 depth, random camera pose deltas)"

  def __init__(self, args):
    self.B, self.H, self.W = 256, 480, 640
    self.scene = args.scene
    self.units_per_step = self.B
    self.algo_bytes_per_step = self.B * self.H * self.W * 12  # 4 B in, 8 B out per pixel
    self.name = "BASELINE config 3: camera_affine_grid on 256 x 480x640 depth frames, random pose deltas, per GPU"

  def config(self, world):
    return {"workload": self.name, "scene": self.scene, "frames_per_step_per_gpu": self.B,
            "parallelism": f"batch-sharded x{world}, no collective",
            "l2": "0.94 GB moved per step, larger than the 126 MB L2; no flush needed"}

  def _deltas(self, n, seed, device=None):
    from dungeon_maps_b200 import synth
    u = synth.uniform((n, 3), seed ^ 0xF10, -1.0, 1.0, device)
    return u * torch.tensor([0.25, 0.25, 0.3], dtype=torch.float32, device=device)

  def setup(self, dev, rank):
    import dungeon_maps_b200 as dmap
    from dungeon_maps_b200 import synth
    self.dev = dev
    self.depth, _, _ = synth.frames(self.scene, self.B, self.H, self.W, 0, seed=rank, device=dev)
    self.delta_host = self._deltas(self.B, rank).cpu()
    self.proj = dmap.MapProjector(width=self.W, height=self.H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0.,
                                  height_offset=0., cam_pitch=PITCH, cam_height=CAM_H, map_res=RES, map_width=MW,
                                  map_height=MH, device=dev)

  def step(self):
    self.out = self.proj.camera_affine_grid(self.depth, self.delta_host)
    return self.out

  def e2e_setup(self):
    self.h_depth = self.depth.cpu().pin_memory()
    self.o_grid = torch.empty((self.B, 1, self.H, self.W, 2), dtype=torch.float32).pin_memory()
    self.h2d = self.h_depth.numel() * 4 + self.B * 192
    self.d2h = self.o_grid.numel() * 4
    self.e2e_path = ("hostapi.camera_affine_grid_host -> MapProjector.camera_affine_grid on pinned host tensors "
                     "(32-frame chunks: H2D copy, kernel, D2H copy of the grid on three streams)")

  def e2e_step(self):
    from dungeon_maps_b200 import hostapi
    hostapi.camera_affine_grid_host(self.proj, self.h_depth, self.delta_host, out=self.o_grid)

  def e2e_check(self):
    assert torch.equal(self.o_grid.nan_to_num(), self.out.cpu().nan_to_num()), "host path and device path disagree"

  def cpu_setup(self, threads):
    from dungeon_maps_b200 import synth
    from oracle import dm_oracle as orc
    n = min(self.B, max(4 * threads, 16))
    depth, _, _ = synth.frames(self.scene, n, self.H, self.W, 0, seed=0)
    d, p = depth.numpy(), self._deltas(n, 0).numpy()
    fx, fy, cx, cy = intrinsics(self.W, self.H)
    self.cpu_units = n
    self.cpu_fn = lambda: orc.camera_affine_grid(d, p, PITCH, CAM_H, fx, fy, cx, cy, threads=threads)
    self.cpu_sample = f"{n} frames per pass ({self.scene} scene), oracle/dm_oracle.c with {threads} OpenMP threads"


class BuilderWorkload:
  """BASELINE config 4: MapBuilder global fusion, 32 environments, 400x400 local height maps merged into
  the world maps over a 100-step walk.  key='builder': the reference's merge (fuse_topdown_maps: re-scatter
  into a freshly sized canvas every step, one host sync).  key='builder_fixed': opt-in in-place merge
  into fixed 2400x2400 canvases (no reference equivalent)."""
  unit = "env-steps/s"
  EPISODE = 100
  e2e_min_steps = 50

  def __init__(self, args, key="builder"):
    self.key = key
    self.B, self.H, self.W = 32, 480, 640
    self.units_per_step = self.B
    self.fixed = key == "builder_fixed"
    self.metric = ("MapBuilder env-steps/sec (32 envs, 480x640 depth -> 400x400 local height map -> merged into "
                   + ("fixed 2400x2400 world maps in place)" if self.fixed else "the growing world map, reference semantics)"))
    self.kernel = ("dm::fuse_inplace_kernel" if self.fixed else
                   "dm_fuse_bbox_i64 + dm_fuse_scatter_f32 (merge kernels; roofline over merge time only)")
    self.name = ("BASELINE config 4: MapBuilder.step over 32 envs x 100-step walk, " +
                 ("fixed 2400x2400 canvases, in-place max-merge" if self.fixed else
                  "fuse_topdown_maps semantics (data-dependent canvas, batch-wide bounding box)"))
    self.t = 0
    self.merge_ms_events = []
    self.algo_bytes_total = 0

  def config(self, world):
    return {"workload": self.name, "scene": "room (66 m hall, 32 walkers)", "envs_per_gpu": self.B,
            "episode_steps": self.EPISODE, "world_cells_at_end": getattr(self, "world_shape", None),
            "parallelism": f"environment-sharded x{world}, no collective",
            "l2": "world maps (>1 GB per step) are larger than the 126 MB L2; no flush needed"}

  @staticmethod
  def walk(b, steps, seed, half=30.0):
    """(steps, b, 3) poses: discrete walk like the reference sim (sim/dungeon.py:244-255): forward 0.25 m / turn ±30°."""
    from dungeon_maps_b200 import synth
    pose = synth.poses(b, seed, xz=half).numpy().astype(np.float64)
    act = (synth.hash_u24(steps * b, seed ^ 0xAC7).numpy() % 4).reshape(steps, b)
    out = np.zeros((steps, b, 3), np.float32)
    for t in range(steps):
      out[t] = pose
      fwd = act[t] < 2
      pose[:, 0] += np.where(fwd, 0.25 * np.sin(pose[:, 2]), 0.0)
      pose[:, 1] += np.where(fwd, 0.25 * np.cos(pose[:, 2]), 0.0)
      pose[:, 2] += np.where(act[t] == 2, math.radians(30), 0.0) - np.where(act[t] == 3, math.radians(30), 0.0)
      pose[:, :2] = np.clip(pose[:, :2], -half, half)
    return torch.from_numpy(out)

  def frames_for(self, poses, dev, seed):
    from dungeon_maps_b200 import synth
    return [synth.room_depth(poses.shape[1], self.H, self.W, HFOV, PITCH, CAM_H, poses[t] if dev is None else poses[t].to(dev), seed, half=33.0,
                             device=dev) for t in range(poses.shape[0])]

  def make_builder(self, dev):
    import dungeon_maps_b200 as dmap
    proj = dmap.MapProjector(width=self.W, height=self.H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0.,
                             height_offset=0., cam_pitch=PITCH, cam_height=CAM_H, map_res=RES, map_width=MW,
                             map_height=MH, trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10,
                             fill_value=dmap.NINF, to_global=True, device=dev)
    if self.fixed:
      return dmap.MapBuilder(map_projector=proj, fixed_canvas=(2400, 2400))
    return dmap.MapBuilder(map_projector=proj)

  def setup(self, dev, rank):
    self.dev = dev
    self.poses = self.walk(self.B, self.EPISODE, rank)
    self.frames = self.frames_for(self.poses, dev, rank)
    self.builder = self.make_builder(dev)
    self.stream = torch.cuda.current_stream(dev)
    self.local_kw = dict(to_global=False, width_offset=MW / 2., height_offset=0., map_width=MW, map_height=MH)

  def _account(self, before, local, after):
    cells = lambda m: 0 if m is None or m.is_empty else m.mask.numel()
    # height-map merge: read (height f32 + mask u8) of both sources, write (height f32 + mask u8) of the new world
    self.algo_bytes_total += 5 * (cells(before) + cells(local)) + 5 * cells(after)

  def step(self):
    t = self.t % self.EPISODE
    if t == 0:
      self.builder.reset()
    local = self.builder.plot(self.frames[t], cam_pose=self.poses[t], **self.local_kw)
    before = self.builder.world_map
    # (the stream object is looked up once: Event.record() without one costs a torch.cuda.current_stream() each time)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(self.stream)
    self.builder.merge(local, keep_pose=False)
    e1.record(self.stream)
    after = self.builder.world_map
    if self.fixed:
      self.algo_bytes_total += 5 * local.mask.numel() + 8 * int(local.mask.shape[0]) * MH * MW
    else:
      self._account(before, local, after)
    self.merge_ms_events.append((e0, e1))
    self.world_shape = list(after.mask.shape)
    self.t += 1
    self.out = after
    return after

  def reset_counters(self):
    self.merge_ms_events, self.algo_bytes_total = [], 0

  def merge_ms(self):
    return sum(a.elapsed_time(b) for a, b in self.merge_ms_events)

  def e2e_setup(self):
    self.h_frames = [f.cpu().pin_memory() for f in self.frames[:10]]
    self.e2e_builder = self.make_builder(self.dev)
    self.h2d = self.h_frames[0].numel() * 4 + self.B * 192
    self.d2h = 40  # the merge's bounding box; the world map stays on the device, like the reference's
    self.e2e_t = 0
    self.e2e_path = "MapBuilder.step on pinned host depth (H2D copy, plot, merge incl. its bbox D2H sync)"

  def e2e_step(self):
    t = self.e2e_t % len(self.h_frames)
    if t == 0:
      self.e2e_builder.reset()
    d = self.h_frames[t].to(self.dev, non_blocking=True)
    self.e2e_builder.step(d, cam_pose=self.poses[t], **self.local_kw)
    torch.cuda.current_stream(self.dev).synchronize()
    self.e2e_t += 1

  def e2e_check(self):
    pass

  def cpu_setup(self, threads):
    from oracle import dm_oracle as orc
    n, T = 2, 4
    poses = self.walk(n, T, 0)
    frames = [f.numpy() for f in self.frames_for(poses, None, 0)]
    fx, fy, cx, cy = intrinsics(self.W, self.H)

    def episode():
      world = None
      for t in range(T):
        p = poses[t].numpy()
        top, mask, hgt = orc.orth_project(frames[t], None, None, p, MW / 2., 0., PITCH, CAM_H, RES, MW, MH, fx, fy, cx,
                                          cy, 0.15, 5.05, None, 10, False, True, -np.inf, None, True, threads=threads)
        src = [orc.FuseSource(hgt, mask, None, MW / 2., 0., RES, True, False, p)]
        if world is not None:
          src.insert(0, orc.FuseSource(world["height"], world["mask"], None, world["width_offset"],
                                       world["height_offset"], RES, True, True, p))
        world = orc.fuse(src, True, p, RES, True) or world
    self.cpu_units = n * T
    self.cpu_fn = episode
    self.cpu_sample = (f"{n} envs x {T} steps per pass (plot via oracle/dm_oracle.c with {threads} threads, merge via the "
                       f"numpy restatement of fuse_topdown_maps)")


def make_workload(args):
  if args.workload in ("proj", "proj5", "proj_labels", "proj5_labels"):
    return ProjWorkload(args, args.workload)
  if args.workload == "flow":
    return FlowWorkload(args)
  return BuilderWorkload(args, args.workload)


# ================================================================================================

_RESULT_LINE = []


def _emit(line: str) -> None:
  """The JSON line of this process (printed by main() once fd 1 is the real stdout again)."""
  _RESULT_LINE[:] = [line]


def cpu_baseline(wl, threads: int, budget_s: float = 20.0):
  """The oracle port on the same workload: bounded sample (≈10-20 s of CPU work)."""
  wl.cpu_setup(threads)
  done, t0 = 0, time.perf_counter()
  while True:
    wl.cpu_fn()
    done += wl.cpu_units
    el = time.perf_counter() - t0
    if el > budget_s / 2 or (el >= 5.0 and done >= 8 * max(wl.units_per_step, wl.cpu_units)):
      break
  return {"value": done / el, "unit": wl.unit, "cores": threads, "kind": "port",
          "sample": f"{wl.cpu_sample}; {done} units in {el:.1f} s"}


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  wl = make_workload(args)
  if wl.key == "builder_fixed":
    _emit(json.dumps({"impl": "reference", "unavailable": "the fixed-canvas merge has no reference implementation"}))
    return
  threads = os.cpu_count() or 1
  wl.cpu_setup(threads)
  for _ in range(max(args.warmup, 1)):
    wl.cpu_fn()
  t0 = time.perf_counter()
  wl.cpu_fn()
  one = time.perf_counter() - t0
  steps = max(3, min(args.steps, int(120.0 / max(one, 1e-3))))  # keep the whole run within minutes
  t0 = time.perf_counter()
  for _ in range(steps):
    wl.cpu_fn()
  el = time.perf_counter() - t0
  value = wl.cpu_units * steps / el
  _emit(json.dumps({
    "impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus, "steps": steps,
    "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": dict(wl.config(1), units_per_step=wl.cpu_units),
    "cpu_baseline": {"value": value, "unit": wl.unit, "cores": threads, "kind": "port", "sample": wl.cpu_sample},
    "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }))


def run_ours(args):
  import torch.distributed as dist
  from dungeon_maps_b200 import _native as nat, shard

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line and nothing else
    dist.init_process_group("nccl", device_id=dev)
    shard.bind_to_gpu_numa_node(local)  # the e2e leg streams GBs through pinned host memory per rank
  barrier = lambda: shard.barrier(dev)
  max_over_ranks = lambda x: shard.max_over_ranks(x, dev)
  sum_over_ranks = lambda x: shard.sum_over_ranks(x, dev)

  wl = make_workload(args)
  wl.setup(dev, rank)  # every rank owns its own environments (weak scaling, no data-path collective)
  warmup = max(args.warmup, 3)
  for _ in range(warmup):
    wl.step()
  if hasattr(wl, "reset_counters"):
    if isinstance(wl, BuilderWorkload):
      wl.t = 0
    wl.reset_counters()
  barrier()
  sampler = ClockSampler(local) if rank == 0 else None
  if sampler:
    sampler.wait_ready()
  barrier()
  launches0 = nat.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t_wall0 = time.perf_counter()
  ev0.record()
  for _ in range(args.steps):
    wl.step()
  ev1.record()
  barrier()
  t_wall1 = time.perf_counter()
  launches = nat.launch_count() - launches0
  ms_total = max_over_ranks(ev0.elapsed_time(ev1))
  clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
  ms_per_step = ms_total / args.steps
  value = world * wl.units_per_step * args.steps / (ms_total * 1e-3)
  total_launches = int(sum_over_ranks(float(launches)))
  if isinstance(wl, BuilderWorkload):  # roofline of the merge kernels over the merge time only
    kernel_ms = max_over_ranks(wl.merge_ms()) / args.steps
    algo_bytes = wl.algo_bytes_total / args.steps
  else:
    kernel_ms, algo_bytes = ms_per_step, wl.algo_bytes_per_step

  # ---- e2e: HOST buffers through the public entry, copies inside the timed region
  # cheap steps (a MapBuilder step is < 1 ms) are timed over more of them: five would be one hiccup away from noise
  e2e_steps = max(2, min(args.steps, max(args.e2e_steps, getattr(wl, "e2e_min_steps", 0))))
  e2e_error = None
  try:  # the leg pins GBs of host memory per rank: a box that refuses must not cost the device-resident line
    wl.e2e_setup()
    wl.e2e_step()
    wl.e2e_step()
  except (RuntimeError, MemoryError) as e:
    e2e_error = f"{type(e).__name__}: {e}"[:200]
  # every rank takes the same branch (a collective below): one failing rank cancels the leg for all
  if sum_over_ranks(1.0 if e2e_error else 0.0) > 0:
    e2e_error = e2e_error or "another rank could not set the host-buffer leg up"
    e2e_value = None
    wl.h2d = getattr(wl, "h2d", 0)
    wl.d2h = getattr(wl, "d2h", 0)
    wl.e2e_path = getattr(wl, "e2e_path", "") + " [not measured: " + e2e_error + "]"
  else:
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      wl.e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * wl.units_per_step * e2e_steps / e2e_s
    wl.e2e_check()

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu = cpu_baseline(wl, os.cpu_count() or 1)

  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  if rank != 0:
    return
  peak, peak_src = measured_peak()
  achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
  roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
              "traffic": ncu_traffic(wl.key), "peak_source": peak_src, "algorithmic_bytes_per_step": algo_bytes,
              "kernel": wl.kernel}
  if isinstance(wl, BuilderWorkload):
    roofline["kernel_ms_per_step"] = kernel_ms
  _emit(json.dumps({
    "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps,
    "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": wl.config(world),
    "clocks": clocks,
    "e2e": {"value": e2e_value, "unit": wl.unit, "h2d_bytes_per_step": int(wl.h2d), "d2h_bytes_per_step": int(wl.d2h),
            "steps": e2e_steps, "path": wl.e2e_path},
    "gpu_launches": total_launches,
    "roofline": roofline,
    "cpu_baseline": cpu,
  }))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=10)
  ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
  ap.add_argument("--workload", choices=("proj", "flow", "builder", "builder_fixed", "proj5", "proj_labels", "proj5_labels"),
                  default="proj")
  ap.add_argument("--scene", choices=("room", "iid"), default="room")
  ap.add_argument("--e2e-steps", type=int, default=5)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  args = ap.parse_args()
  if args.gpus > 1 and args.impl == "ours" and "WORLD_SIZE" not in os.environ:
    # called directly with --gpus N: become the torchrun launch the driver would have made (one rank per GPU)
    import socket
    with socket.socket() as sock:
      sock.bind(("127.0.0.1", 0))
      port = sock.getsockname()[1]
    os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                              f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", str(port),
                              os.path.abspath(__file__)] + sys.argv[1:])
  # stdout carries the one JSON line and nothing else: libraries that print from C (NCCL's version banner ignores
  # NCCL_DEBUG_FILE on some boxes) get stderr as their fd 1 while the run lasts
  sys.stdout.flush()
  real_stdout = os.dup(1)
  os.dup2(2, 1)
  try:
    if args.impl == "reference":
      run_reference(args)
    else:
      run_ours(args)
  finally:
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
  if _RESULT_LINE:
    print(_RESULT_LINE[0], flush=True)


if __name__ == "__main__":
  main()
