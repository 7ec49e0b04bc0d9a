#!/usr/bin/env python
"""bench.py — top-down maps/sec of the fused projection (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scene room|iid] [--impl ours|reference]

One "step" = one pass of the hot path (dm_orth_project_f32: fused projection + resolve) over one
batch of 64 synthetic 480x640 depth frames + 16 one-hot semantic channels → 400x400 local maps
(topdown, mask, height).  N > 1: one process per GPU (torchrun), every rank projects its own
64 environments, no collective on the data path (weak scaling).  Prints ONE JSON line on rank 0.
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

B, H, W, C, MH, MW = 64, 480, 640, 16, 400, 400
HFOV, PITCH, CAM_H, RES = math.radians(70), math.radians(-10), 0.88, 0.03
METRIC = "top-down maps/sec (batch 64, 640x480 depth + 16 semantic channels -> 400x400 maps)"
UNIT = "maps/s"
# read 4*N*(1+C) input bytes, write Mh*Mw*(4C + C + 4) output bytes per frame
ALGO_BYTES_PER_FRAME = 4 * H * W * (1 + C) + MH * MW * (4 * C + C + 4)


def proj_kwargs():
  cx, cy = W / 2., H / 2.
  fx = cx / np.tan(HFOV / 2.)
  return dict(map_res=RES, map_width=MW, map_height=MH, focal_x=fx, focal_y=fx, center_x=cx, center_y=cy,
              trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None, clip_border=10,
              to_global=False, flip_h=True, fill_value=-np.inf, reduction=None, get_height_map=True)


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


def ncu_traffic():
  """dram bytes per launch of the dominant kernel from the committed ncu --set full summary."""
  try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
      return json.load(f).get("dram_bytes_per_step")
  except Exception:
    return None


def cpu_baseline(scene: str, threads: int, budget_s: float = 20.0):
  """The oracle port (scalar C restatement of the reference, OpenMP over frames) on the same workload."""
  from dungeon_maps_b200 import synth
  from oracle import dm_oracle as orc
  n = min(B, max(threads, 8))
  depth, values, pose = synth.frames(scene, n, H, W, C, seed=0)
  d, v, p = depth.numpy(), values.numpy(), pose.numpy()
  kw = proj_kwargs()
  orc.orth_project(d[:1], v[:1], None, p[:1], 200., 0., PITCH, CAM_H, threads=1, **kw)  # warm-up / page-in
  done, t0 = 0, time.perf_counter()
  while True:
    orc.orth_project(d, v, None, p, 200., 0., PITCH, CAM_H, threads=threads, **kw)
    done += n
    el = time.perf_counter() - t0
    if el > budget_s / 2 or done >= 8 * B:
      break
  return {"value": done / el, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": f"{done} frames of the same workload ({scene} scene), oracle/dm_oracle.c with {threads} OpenMP threads, {el:.1f} s"}


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  from dungeon_maps_b200 import synth
  from oracle import dm_oracle as orc
  threads = os.cpu_count() or 1
  n = min(B, max(threads, 8))  # bounded sample: one frame per host thread, at most the batch
  depth, values, pose = synth.frames(args.scene, n, H, W, C, seed=0)
  d, v, p = depth.numpy(), values.numpy(), pose.numpy()
  kw = proj_kwargs()
  step = lambda: orc.orth_project(d, v, None, p, 200., 0., PITCH, CAM_H, threads=threads, **kw)
  for _ in range(max(args.warmup, 1)):
    step()
  steps = args.steps
  t0 = time.perf_counter()
  step()
  one = time.perf_counter() - t0
  steps = max(3, min(steps, int(120.0 / max(one, 1e-3))))  # keep the whole run within minutes
  t0 = time.perf_counter()
  for _ in range(steps):
    step()
  el = time.perf_counter() - t0
  value = n * steps / el
  sample = f"{n} frames per step ({args.scene} scene), oracle port of the reference CPU path, {threads} OpenMP threads"
  print(json.dumps({
    "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
    "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": {"workload": "BASELINE config 2: 64 x 480x640 depth + 16-channel one-hot semantics -> 400x400 maps "
                           "(topdown + mask + height)", "scene": args.scene, "frames_per_step": n},
    "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }))


def run_ours(args):
  import torch.distributed as dist
  import dungeon_maps_b200 as dmap
  from dungeon_maps_b200 import _native as nat, hostapi, synth

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

  from dungeon_maps_b200 import shard
  barrier = lambda: shard.barrier(dev)
  max_over_ranks = lambda x: shard.max_over_ranks(x, dev)
  sum_over_ranks = lambda x: shard.sum_over_ranks(x, dev)

  kw = proj_kwargs()
  # every rank owns 64 environments (weak scaling, no data-path collective)
  depth, values, pose = synth.frames(args.scene, B, H, W, C, seed=rank, device=dev)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=200., height_offset=0.,
                           cam_pitch=PITCH, cam_height=CAM_H, map_res=RES, map_width=MW, map_height=MH,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10, to_global=False,
                           fill_value=dmap.NINF, device=dev)
  pose_host = pose.cpu()

  def step():
    return proj.orth_project(depth, values, cam_pose=pose_host, get_height_map=True)

  for _ in range(max(args.warmup, 3)):
    out = step()
  barrier()
  sampler = ClockSampler(local) if rank == 0 else None
  launches0 = nat.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t_wall0 = time.perf_counter()
  ev0.record()
  for _ in range(args.steps):
    out = step()
  ev1.record()
  barrier()
  t_wall1 = time.perf_counter()
  launches = nat.launch_count() - launches0
  ms_total = max_over_ranks(ev0.elapsed_time(ev1))
  clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
  ms_per_step = ms_total / args.steps
  value = world * B * args.steps / (ms_total * 1e-3)
  total_launches = int(sum_over_ranks(float(launches)))

  # ---- e2e: HOST buffers through the C ABI (dm_orth_project_host_f32), copies inside the timed region
  e2e_steps = max(2, min(args.steps, args.e2e_steps))
  h_depth = depth.cpu().pin_memory().numpy()
  h_values = values.cpu().pin_memory().numpy()
  o_top = torch.empty((B, C, MH, MW), dtype=torch.float32).pin_memory().numpy()
  o_mask = torch.empty((B, C, MH, MW), dtype=torch.uint8).pin_memory().numpy()
  o_hgt = torch.empty((B, 1, MH, MW), dtype=torch.float32).pin_memory().numpy()
  host_step = lambda: hostapi.orth_project_host(h_depth, h_values, None, pose_host, 200., 0., PITCH, CAM_H,
                                                device=local, out=(o_top, o_mask, o_hgt), **kw)
  host_step()
  host_step()
  barrier()
  t0 = time.perf_counter()
  for _ in range(e2e_steps):
    host_step()
  torch.cuda.synchronize(dev)
  e2e_s = max_over_ranks(time.perf_counter() - t0)
  e2e_value = world * B * e2e_steps / e2e_s
  h2d = h_depth.nbytes + h_values.nbytes + B * 192
  d2h = o_top.nbytes + o_mask.nbytes + o_hgt.nbytes
  assert np.array_equal(o_top, out[0].cpu().numpy()) and np.array_equal(o_hgt, out[2][:, :1].cpu().numpy()), \
      "host-buffer path and device path disagree"

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu = cpu_baseline(args.scene, os.cpu_count() or 1)

  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  if rank != 0:
    return
  peak, peak_src = measured_peak()
  algo_bytes = ALGO_BYTES_PER_FRAME * B
  achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
  print(json.dumps({
    "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
    "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": {"workload": "BASELINE config 2: 64 x 480x640 depth + 16-channel one-hot semantics -> 400x400 maps "
                           "(topdown + mask + height), per GPU", "scene": args.scene, "fill_value": "-inf",
               "frames_per_step_per_gpu": B, "parallelism": f"batch-sharded x{world}, no collective",
               "l2": "inputs (1.34 GB per step) are larger than the 126 MB L2; no flush needed"},
    "clocks": clocks,
    "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": e2e_steps, "path": "hostapi.orth_project_host -> dm_orth_project_host_f32 (pinned host buffers)"},
    "gpu_launches": total_launches,
    "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                 "traffic": ncu_traffic(), "peak_source": peak_src,
                 "algorithmic_bytes_per_step": algo_bytes,
                 "kernel": "dm::proj_ws_kernel (one step = ONE persistent launch of dm_orth_project_f32: projection + resolve)"},
    "cpu_baseline": cpu,
  }))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=10)
  ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
  ap.add_argument("--scene", choices=("room", "iid"), default="room")
  ap.add_argument("--e2e-steps", type=int, default=5)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
