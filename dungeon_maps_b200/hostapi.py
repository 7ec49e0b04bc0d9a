"""HOST-buffer entry of the projection: numpy / pinned CPU tensors in, numpy out, through
`dm_orth_project_host_f32` (host→device copy, kernels, device→host copy as a three-stream
pipeline inside the library).  This is what a non-torch host (or the reference's numpy-facing
MapBuilder.step) would call, and what bench.py times as `e2e`.
"""
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from . import _params as prm
from . import utils


def _host(a, dtype):
  if a is None:
    return None
  if torch.is_tensor(a):
    a = a.detach().cpu().numpy()
  return np.ascontiguousarray(a, dtype=dtype)


def _batch_of(a: np.ndarray, b: int, what: str) -> np.ndarray:
  """A batch-1 image array is shared by all b frames (the kernels index every plane by frame)."""
  if a.shape[0] == b:
    return a
  if a.shape[0] == 1:
    return np.ascontiguousarray(np.broadcast_to(a, (b,) + a.shape[1:]))
  raise RuntimeError(f"{what} has batch {a.shape[0]}, depth_map has batch {b}")


def orth_project_host(depth_map, value_map, valid_map, cam_pose, width_offset, height_offset, cam_pitch,
                      cam_height, map_res, map_width, map_height, focal_x, focal_y, center_x, center_y,
                      trunc_depth_min, trunc_depth_max, trunc_height_max, clip_border, to_global, flip_h=True,
                      fill_value=None, reduction=None, get_height_map=False, device: Optional[int] = None, out=None,
                      label_map=None, num_classes: Optional[int] = None):
  """Same arguments as maps.orth_project with (b,1,H,W) / (b,C,H,W) HOST arrays; returns numpy
  (topdown, mask[, height]).  `out` may carry preallocated (pinned) result arrays.  `label_map` (b,1,H,W) uint8
  class ids + `num_classes` instead of `value_map`: the result of value_map = one_hot(label_map) with 5 instead
  of 4 * (C + 1) bytes per pixel crossing PCIe (dm_orth_project_labels_host_f32).  `device` defaults to the
  current CUDA device; the caller's current device is left as it was."""
  device = nat.require_cuda(device).index
  if utils._reduction_code(reduction, fused=False) > 1:
    # sum / mean / prod: the composed path on device tensors (maps.orth_project), results copied back
    from . import maps
    dev = torch.device("cuda", device)
    up = lambda a, dt: None if a is None else torch.as_tensor(np.asarray(a)).to(device=dev, dtype=dt)
    res = maps.orth_project(up(depth_map, torch.float32), up(value_map, torch.float32), up(valid_map, torch.bool),
                            cam_pose, width_offset, height_offset, cam_pitch, cam_height, map_res, map_width,
                            map_height, focal_x, focal_y, center_x, center_y, trunc_depth_min, trunc_depth_max,
                            trunc_height_max, clip_border, to_global, flip_h, fill_value, reduction, get_height_map,
                            device=dev, label_map=None if label_map is None else up(label_map, torch.int64),
                            num_classes=num_classes)
    return tuple(r.contiguous().cpu().numpy() for r in res)
  depth = _host(depth_map, np.float32)
  values = _host(value_map, np.float32)
  valid = None if valid_map is None else _host(np.asarray(valid_map).astype(bool), np.uint8)
  b, _, H, W = depth.shape
  labels = None
  if label_map is not None:
    if values is not None or num_classes is None or not 1 <= int(num_classes) <= 63:
      raise ValueError("label_map excludes value_map and needs num_classes in 1..63")
    labels = label_map.detach().cpu().numpy() if torch.is_tensor(label_map) else np.asarray(label_map)
    if labels.dtype != np.uint8:
      if labels.dtype.kind not in "iu":
        raise TypeError(f"label_map must hold integer class ids, got {labels.dtype}")
      labels = np.where((labels < 0) | (labels >= int(num_classes)), 255, labels).astype(np.uint8)
    labels = _batch_of(np.ascontiguousarray(labels).reshape((-1, 1, H, W)), b, "label_map")
  if values is not None:
    values = _batch_of(values, b, "value_map")
  if valid is not None:
    valid = _batch_of(valid.reshape((-1, 1, H, W)), b, "valid_map")
  C = int(num_classes) if labels is not None else (0 if values is None else values.shape[1])
  n_points = H * W
  pose = prm.per_sample(cam_pose, b, (3,), "cam_pose")
  pitch, camh = prm.per_sample(cam_pitch, b), prm.per_sample(cam_height, b)
  samples = torch.zeros((b, nat.PROJ_SAMPLE_WORDS), dtype=torch.float32)
  samples[:, 0:16] = prm.camera_to_local(pitch, camh, n_points)
  samples[:, 16:32] = prm.local_to_global(pose, n_points) if to_global else prm.identity(b)
  samples[:, 32] = prm.per_sample(width_offset, b)
  samples[:, 33] = prm.per_sample(height_offset, b)
  samples = samples.numpy()
  cfg = nat.DmProjCfg()
  cfg.H, cfg.W, cfg.C, cfg.Mh, cfg.Mw = H, W, C, int(map_height), int(map_width)
  cfg.fx, cfg.fy, cfg.cx, cfg.cy = focal_x, focal_y, center_x, center_y
  cfg.map_res = map_res
  cfg.has_trunc_depth_min = trunc_depth_min is not None
  cfg.has_trunc_depth_max = trunc_depth_max is not None
  cfg.has_trunc_height_max = trunc_height_max is not None
  cfg.trunc_depth_min = trunc_depth_min or 0.
  cfg.trunc_depth_max = trunc_depth_max or 0.
  cfg.trunc_height_max = trunc_height_max or 0.
  cfg.clip_border = int(clip_border) if clip_border is not None else 0
  cfg.flip_h = bool(flip_h)
  cfg.fill_value = 0. if fill_value is None else fill_value
  want_h = bool(get_height_map) and C > 0
  cfg.want_height = want_h
  cfg.reduction = utils._reduction_code(reduction)
  cfg.fast_steps = prm.fast_steps(torch.from_numpy(samples))
  Cv = max(C, 1)
  if out is None:
    top = np.empty((b, Cv, cfg.Mh, cfg.Mw), np.float32)
    mask = np.empty((b, Cv, cfg.Mh, cfg.Mw), np.uint8)
    hgt = np.empty((b, 1, cfg.Mh, cfg.Mw), np.float32) if want_h else None
  else:
    top, mask, hgt = out
  p = lambda a: None if a is None else a.ctypes.data
  if labels is not None:
    rc = nat.lib().dm_orth_project_labels_host_f32(p(depth), p(labels), p(valid), p(samples), cfg, b, p(top), p(mask),
                                                   p(hgt), int(device))
    nat.check(rc, "dm_orth_project_labels_host_f32")
  else:
    rc = nat.lib().dm_orth_project_host_f32(p(depth), p(values), p(valid), p(samples), cfg, b, p(top), p(mask),
                                            p(hgt), int(device))
    nat.check(rc, "dm_orth_project_host_f32")
  mask_b = mask.view(np.bool_)
  if not get_height_map:
    return top, mask_b
  return top, mask_b, (top if C == 0 else hgt)


_flow_streams = {}


def camera_affine_grid_host(proj, depth_map: torch.Tensor, trans_pose, out: Optional[torch.Tensor] = None,
                            chunk: int = 32, **kwargs) -> torch.Tensor:
  """MapProjector.camera_affine_grid (maps.py:353-460) for HOST depth maps: the batch goes through the device in
  chunks of `chunk` frames — host→device copy, kernel, device→host copy of the grid on three streams — so that both
  PCIe directions and the kernel overlap (the grid is twice the size of the depth: one blocking round trip is
  copy-out bound and leaves the copy-in engine idle two thirds of the time).
  depth_map: (b, 1, H, W) float32 CPU tensor (pinned for full-rate copies); trans_pose: (b, 3) or (3,);
  out: optional (b, 1, H, W, 2) float32 CPU tensor (pinned).  Returns `out`."""
  dev = nat.require_cuda(proj.device)
  depth = depth_map if torch.is_tensor(depth_map) else torch.as_tensor(np.asarray(depth_map, dtype=np.float32))
  assert depth.dim() == 4 and depth.shape[1] == 1 and depth.dtype is torch.float32 and depth.device.type == "cpu", \
      "depth_map must be a (b, 1, H, W) float32 CPU tensor"
  b, _, H, W = depth.shape
  pose = prm.per_sample(trans_pose, b, (3,), "trans_pose")
  if out is None:
    out = torch.empty((b, 1, H, W, 2), dtype=torch.float32).pin_memory()
  streams = _flow_streams.get(dev.index)
  if streams is None:
    streams = _flow_streams[dev.index] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
  s_in, s_out = streams
  cur = torch.cuda.current_stream(dev)
  s_in.wait_stream(cur)  # nothing the caller queued is overtaken
  for i0 in range(0, b, max(int(chunk), 1)):
    i1 = min(b, i0 + max(int(chunk), 1))
    with torch.cuda.stream(s_in):
      d = depth[i0:i1].to(dev, non_blocking=True)
    d.record_stream(cur)
    cur.wait_stream(s_in)
    grid = proj.camera_affine_grid(d, pose[i0:i1], **kwargs)
    grid.record_stream(s_out)
    s_out.wait_stream(cur)
    with torch.cuda.stream(s_out):
      out[i0:i1].copy_(grid, non_blocking=True)
  s_out.synchronize()
  return out
