"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/dm_oracle.c for the header).

Python face of the scalar C restatement: builds the per-sample parameter
blocks the way the reference's host code would (Rodrigues matrices with the
reference's own torch-CPU op sequence, utils.py:303-327) and calls
libdm_oracle.so through ctypes.  numpy in, numpy out, batch > 1 supported as
"reference applied per sample" (SURVEY.md D6).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import ctypes
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import build as _build

ANGLE_EPS = 0.001  # utils.py:47
NINF = -np.inf

STEP_DT = np.dtype([("R", "<f4", (9,)), ("t", "<f4", (3,)), ("kind", "<i4"), ("fused", "<i4"), ("_pad", "<i4", (2,))])
PROJ_SAMPLE_DT = np.dtype([
  ("to_local", STEP_DT), ("to_global", STEP_DT),
  ("width_offset", "<f4"), ("height_offset", "<f4"), ("_pad", "<f4", (14,)),
])
FLOW_SAMPLE_DT = np.dtype([("to_local", STEP_DT), ("transition", STEP_DT), ("to_camera", STEP_DT)])
assert STEP_DT.itemsize == 64 and PROJ_SAMPLE_DT.itemsize == 192 and FLOW_SAMPLE_DT.itemsize == 192


class ProjCfg(ctypes.Structure):
  _fields_ = [
    ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("C", ctypes.c_int32),
    ("Mh", ctypes.c_int32), ("Mw", ctypes.c_int32),
    ("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float),
    ("map_res", ctypes.c_float),
    ("trunc_depth_min", ctypes.c_float), ("trunc_depth_max", ctypes.c_float),
    ("trunc_height_max", ctypes.c_float),
    ("has_trunc_depth_min", ctypes.c_int32), ("has_trunc_depth_max", ctypes.c_int32),
    ("has_trunc_height_max", ctypes.c_int32),
    ("clip_border", ctypes.c_int32), ("flip_h", ctypes.c_int32),
    ("fill_value", ctypes.c_float), ("want_height", ctypes.c_int32),
    ("reduction", ctypes.c_int32), ("fast_steps", ctypes.c_int32), ("_pad", ctypes.c_int32 * 3),
  ]


class FlowCfg(ctypes.Structure):
  _fields_ = [
    ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("channels", ctypes.c_int32),
    ("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float),
    ("flip_h", ctypes.c_int32), ("emit_flow", ctypes.c_int32), ("_pad", ctypes.c_int32 * 6),
  ]


_lib = None


def lib() -> ctypes.CDLL:
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(_build.build())
    _lib.dmo_max_threads.restype = ctypes.c_int
  return _lib


def _p(a: Optional[np.ndarray]):
  return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def f32(x) -> np.float32:
  return np.float32(x)


# ---- host-side parameter construction (reference op order) -------------------

def rodrigues(axis: Sequence[float], angle) -> np.ndarray:
  """(b, 9) float32 rotation matrices, utils.py:303-327 op for op (torch CPU)."""
  angle = torch.as_tensor(np.asarray(angle, dtype=np.float32)).reshape(-1, 1)
  batch = angle.shape[0]
  ax = torch.tensor(axis, dtype=torch.float32).view(-1, 3)
  ax = ax / torch.linalg.norm(ax, dim=-1, keepdim=True)
  ax = ax.expand(batch, 3)
  zeros = torch.zeros((batch,), dtype=torch.float32)
  S = torch.stack((zeros, -ax[:, 2], ax[:, 1], ax[:, 2], zeros, -ax[:, 0], -ax[:, 1], ax[:, 0], zeros), dim=-1)
  S3 = S.view(-1, 3, 3)
  S2 = torch.einsum("bij,bjk->bik", S3, S3).reshape(-1, 9)
  eye = torch.eye(3).view(-1, 9)
  angle = torch.where(torch.abs(angle) > ANGLE_EPS, angle, torch.tensor(0.0))
  R = eye + torch.sin(angle) * S + (1 - torch.cos(angle)) * S2
  return R.numpy().astype(np.float32)


def fused_for(n_points: int) -> bool:
  """at::bmm switches from its naive loop to MKL sgemm (FMA) at 9*n >= 400."""
  return 9 * int(n_points) >= 400


def steps(kind: int, R: np.ndarray, t: np.ndarray, n_points: int = 1 << 30) -> np.ndarray:
  """n_points: how many points per sample the reference rotates in this call."""
  out = np.zeros((R.shape[0],), dtype=STEP_DT)
  out["R"] = R
  out["t"] = t
  out["kind"] = kind
  out["fused"] = fused_for(n_points)
  return out


def _vec(x, b) -> np.ndarray:
  a = np.asarray(x, dtype=np.float32).reshape(-1)
  if a.shape[0] == 1 and b > 1:
    a = np.repeat(a, b)
  assert a.shape[0] == b, (a.shape, b)
  return a


def to_local_steps(pitch: np.ndarray, cam_h: np.ndarray, n_points: int = 1 << 30) -> np.ndarray:
  b = pitch.shape[0]
  t = np.zeros((b, 3), np.float32)
  t[:, 1] = cam_h
  return steps(1, rodrigues([1., 0., 0.], pitch), t, n_points)


def to_camera_steps(pitch: np.ndarray, cam_h: np.ndarray, n_points: int = 1 << 30) -> np.ndarray:
  b = pitch.shape[0]
  t = np.zeros((b, 3), np.float32)
  t[:, 1] = -cam_h
  return steps(2, rodrigues([1., 0., 0.], -pitch), t, n_points)


def to_global_steps(pose: np.ndarray, n_points: int = 1 << 30) -> np.ndarray:
  b = pose.shape[0]
  t = np.zeros((b, 3), np.float32)
  t[:, 0] = pose[:, 0]
  t[:, 2] = pose[:, 1]
  return steps(1, rodrigues([0., 1., 0.], pose[:, 2]), t, n_points)


def from_global_steps(pose: np.ndarray, n_points: int = 1 << 30) -> np.ndarray:
  b = pose.shape[0]
  t = np.zeros((b, 3), np.float32)
  t[:, 0] = pose[:, 0]
  t[:, 2] = pose[:, 1]
  return steps(2, rodrigues([0., 1., 0.], -pose[:, 2]), -t, n_points)  # -pos: (-x, -0.0, -z), maps.py:937


def identity_steps(b: int) -> np.ndarray:
  return np.zeros((b,), dtype=STEP_DT)


def intrinsics(width, height, hfov, vfov=None):
  """utils.py:94-116 (float64 host math, cast to f32 where the reference casts)."""
  cx = width / 2.
  cy = height / 2.
  fx = cx / np.tan(hfov / 2.)
  fy = cy / np.tan(vfov / 2.) if vfov is not None else fx
  return dict(cx=cx, cy=cy, fx=fx, fy=fy)


# ---- the path ---------------------------------------------------------------

def orth_project(depth, value_map, valid_map, cam_pose, width_offset, height_offset, cam_pitch,
                 cam_height, map_res, map_width, map_height, focal_x, focal_y, center_x, center_y,
                 trunc_depth_min, trunc_depth_max, trunc_height_max, clip_border, to_global,
                 flip_h=True, fill_value=None, reduction=None, get_height_map=False, threads=1):
  """maps.py:127-351.  depth (b,1,H,W); value_map (b,C,H,W) or None; returns
  (topdown, mask[, height]) as numpy arrays; height is (b,1,Mh,Mw) broadcastable."""
  depth = np.ascontiguousarray(depth, dtype=np.float32)
  assert depth.ndim == 4 and depth.shape[1] == 1
  b, _, H, W = depth.shape
  C = 0
  if value_map is not None:
    value_map = np.ascontiguousarray(value_map, dtype=np.float32)
    C = value_map.shape[1]
  if valid_map is not None:
    valid_map = np.ascontiguousarray(np.asarray(valid_map).astype(bool).astype(np.uint8))
  pose = np.asarray(cam_pose, dtype=np.float32).reshape(-1, 3)
  if pose.shape[0] == 1 and b > 1:
    pose = np.repeat(pose, b, 0)
  samples = np.zeros((b,), dtype=PROJ_SAMPLE_DT)
  samples["to_local"] = to_local_steps(_vec(cam_pitch, b), _vec(cam_height, b), H * W)
  samples["to_global"] = to_global_steps(pose, H * W) if to_global else identity_steps(b)
  samples["width_offset"] = _vec(width_offset, b)
  samples["height_offset"] = _vec(height_offset, b)
  cfg = ProjCfg()
  cfg.H, cfg.W, cfg.C, cfg.Mh, cfg.Mw = H, W, C, int(map_height), int(map_width)
  cfg.fx, cfg.fy, cfg.cx, cfg.cy = f32(focal_x), f32(focal_y), f32(center_x), f32(center_y)
  cfg.map_res = f32(map_res)
  cfg.has_trunc_depth_min = trunc_depth_min is not None
  cfg.has_trunc_depth_max = trunc_depth_max is not None
  cfg.has_trunc_height_max = trunc_height_max is not None
  cfg.trunc_depth_min = f32(trunc_depth_min or 0.)
  cfg.trunc_depth_max = f32(trunc_depth_max or 0.)
  cfg.trunc_height_max = f32(trunc_height_max or 0.)
  cfg.clip_border = int(clip_border or 0)
  cfg.flip_h = bool(flip_h)
  cfg.fill_value = f32(0. if fill_value is None else fill_value)
  cfg.want_height = bool(get_height_map and C > 0)
  red = "max" if reduction is None else str(getattr(reduction, "value", reduction))
  cfg.reduction = {"max": 0, "min": 1}[red]
  Cout = max(C, 1)
  top = np.empty((b, Cout, cfg.Mh, cfg.Mw), np.float32)
  mask = np.empty((b, Cout, cfg.Mh, cfg.Mw), np.uint8)
  hgt = np.empty((b, 1, cfg.Mh, cfg.Mw), np.float32) if cfg.want_height else None
  rc = lib().dmo_orth_project(_p(depth), _p(value_map), _p(valid_map), _p(samples),
                              ctypes.byref(cfg), b, _p(top), _p(mask), _p(hgt), int(threads))
  assert rc == 0
  mask = mask.astype(bool)
  if get_height_map:
    return top, mask, (top if C == 0 else hgt)
  return top, mask


def camera_affine_grid(depth, trans_pose, cam_pitch, cam_height, focal_x, focal_y, center_x,
                       center_y, flip_h=True, emit_flow=False, threads=1):
  """maps.py:353-460.  depth (b,c,H,W) → grid (b,c,H,W,2)."""
  depth = np.ascontiguousarray(depth, dtype=np.float32)
  b, ch, H, W = depth.shape
  pose = np.asarray(trans_pose, dtype=np.float32).reshape(-1, 3)
  if pose.shape[0] == 1 and b > 1:
    pose = np.repeat(pose, b, 0)
  pitch, camh = _vec(cam_pitch, b), _vec(cam_height, b)
  samples = np.zeros((b,), dtype=FLOW_SAMPLE_DT)
  samples["to_local"] = to_local_steps(pitch, camh, ch * H * W)
  samples["transition"] = to_global_steps(pose, ch * H * W)
  samples["to_camera"] = to_camera_steps(pitch, camh, ch * H * W)
  cfg = FlowCfg()
  cfg.H, cfg.W, cfg.channels = H, W, ch
  cfg.fx, cfg.fy, cfg.cx, cfg.cy = f32(focal_x), f32(focal_y), f32(center_x), f32(center_y)
  cfg.flip_h = bool(flip_h)
  cfg.emit_flow = bool(emit_flow)
  grid = np.empty((b, ch, H, W, 2), np.float32)
  rc = lib().dmo_affine_grid(_p(depth), _p(samples), ctypes.byref(cfg), b, _p(grid), int(threads))
  assert rc == 0
  return grid


def transform_points(points, step_list):
  """points (b,n,3); step_list: list of (b,) STEP_DT arrays applied in order."""
  points = np.ascontiguousarray(points, dtype=np.float32)
  b, n, _ = points.shape
  st = np.ascontiguousarray(np.stack(step_list, axis=1))  # (b, k)
  out = np.empty_like(points)
  lib().dmo_transform_points(_p(points), _p(st), st.shape[1], b, ctypes.c_int64(n), _p(out))
  return out


def image_camera(points, fx, fy, cx, cy, flip_h, height, to_image):
  points = np.ascontiguousarray(points, dtype=np.float32)
  flat = points.reshape(-1, 3)
  out = np.empty_like(flat)
  lib().dmo_image_camera(_p(flat), ctypes.c_int64(flat.shape[0]), ctypes.c_float(f32(fx)),
                         ctypes.c_float(f32(fy)), ctypes.c_float(f32(cx)), ctypes.c_float(f32(cy)),
                         int(bool(flip_h)), int(height or 0), int(to_image), _p(out))
  return out.reshape(points.shape)


def depth_to_points(depth, valid, fx, fy, cx, cy, flip_h, tmin, tmax):
  depth = np.ascontiguousarray(depth, dtype=np.float32)
  H, W = depth.shape[-2:]
  frames = int(np.prod(depth.shape[:-2]))
  if valid is not None:
    valid = np.ascontiguousarray(np.broadcast_to(np.asarray(valid).astype(np.uint8), depth.shape))
  pts = np.empty(depth.shape + (3,), np.float32)
  ok = np.empty(depth.shape, np.uint8)
  lib().dmo_depth_to_points(_p(depth), _p(valid), ctypes.c_int64(frames), H, W,
                            ctypes.c_float(f32(fx)), ctypes.c_float(f32(fy)),
                            ctypes.c_float(f32(cx)), ctypes.c_float(f32(cy)), int(bool(flip_h)),
                            int(tmin is not None), ctypes.c_float(f32(tmin or 0.)),
                            int(tmax is not None), ctypes.c_float(f32(tmax or 0.)), _p(pts), _p(ok))
  return pts, ok.astype(bool)


def map_quantize(x, z, woff, hoff, res, map_height, flip_h):
  x = np.ascontiguousarray(x, dtype=np.float32)
  z = np.ascontiguousarray(z, dtype=np.float32)
  b = x.shape[0]
  n = int(np.prod(x.shape[1:]))
  woff, hoff = _vec(woff, b), _vec(hoff, b)
  xb = np.empty(x.shape, np.int64)
  zb = np.empty(x.shape, np.int64)
  lib().dmo_map_quantize(_p(x), _p(z), _p(woff), _p(hoff), b, ctypes.c_int64(n),
                         ctypes.c_float(f32(res)), int(map_height or 0), int(bool(flip_h)),
                         _p(xb), _p(zb))
  return xb, zb


def map_dequantize(xb, zb, woff, hoff, res, map_height, flip_h):
  xb = np.ascontiguousarray(xb, dtype=np.float32)
  zb = np.ascontiguousarray(zb, dtype=np.float32)
  b = xb.shape[0]
  n = int(np.prod(xb.shape[1:]))
  woff, hoff = _vec(woff, b), _vec(hoff, b)
  x = np.empty(xb.shape, np.float32)
  z = np.empty(xb.shape, np.float32)
  lib().dmo_map_dequantize(_p(xb), _p(zb), _p(woff), _p(hoff), b, ctypes.c_int64(n),
                           ctypes.c_float(f32(res)), int(map_height or 0), int(bool(flip_h)),
                           _p(x), _p(z))
  return x, z


def scatter(values, coords, valid, canvas, fill_value, reduction=None):
  """utils.py:389-492 on (..., N) values, (..., N, 2) coords, (..., Mh, Mw) canvas."""
  values = np.ascontiguousarray(values, dtype=np.float32)
  Mh, Mw = canvas.shape[-2:]
  N = values.shape[-1]
  B = int(np.prod(values.shape[:-1]))
  coords = np.ascontiguousarray(np.broadcast_to(coords, values.shape + (2,)), dtype=np.int64)
  if valid is not None:
    valid = np.ascontiguousarray(np.broadcast_to(valid, values.shape).astype(np.uint8))
  cv = np.ascontiguousarray(np.broadcast_to(canvas, values.shape[:-1] + (Mh, Mw)), dtype=np.float32).copy()
  mask = np.empty(cv.shape, np.uint8)
  red = "max" if reduction is None else str(getattr(reduction, "value", reduction))
  rc = lib().dmo_scatter(_p(values), _p(coords), _p(valid), ctypes.c_int64(B), ctypes.c_int64(N),
                         Mh, Mw, int(fill_value is not None),
                         ctypes.c_float(f32(0. if fill_value is None else fill_value)),
                         {"max": 0, "min": 1, "sum": 2, "mean": 3, "prod": 4}[red], _p(cv), _p(mask))
  assert rc == 0
  return cv, mask.astype(bool)


class FuseSource:
  """One TopdownMap handed to fuse_topdown_maps, as plain arrays."""

  def __init__(self, height, mask, values, width_offset, height_offset, map_res, flip_h,
               to_global, cam_pose):
    self.height = np.asarray(height, dtype=np.float32)  # (b,C,h,w), may be a stride-0 view
    self.mask = np.asarray(mask).astype(bool)
    self.values = None if values is None else np.asarray(values, dtype=np.float32)
    self.width_offset = width_offset
    self.height_offset = height_offset
    self.map_res = map_res
    self.flip_h = flip_h
    self.to_global = to_global
    self.cam_pose = cam_pose


def fuse(sources, target_to_global, target_pose, target_res, target_flip_h, fill_value=None,
         proj_fill_value=NINF, reduction=None):
  n_total = sum(int(np.prod(s.mask.shape[1:])) for s in sources)  # points per sample after the cat
  """fuse_topdown_maps, maps.py:2181-2287.  Returns dict(topdown, mask, height,
  map_width, map_height, width_offset, height_offset) or None when no point is valid."""
  pts_x, pts_y, pts_z, masks, vals = [], [], [], [], []
  is_height = all(s.values is None for s in sources)
  for s in sources:
    b, C, h, w = s.mask.shape
    hm = np.broadcast_to(s.height, (b, C, h, w))
    hm_c = np.ascontiguousarray(hm)
    pose = np.asarray(s.cam_pose, np.float32).reshape(-1, 3)
    if pose.shape[0] == 1 and b > 1:
      pose = np.repeat(pose, b, 0)
    tpose = np.asarray(target_pose, np.float32).reshape(-1, 3)
    if tpose.shape[0] == 1 and b > 1:
      tpose = np.repeat(tpose, b, 0)
    st = np.zeros((b, 2), dtype=STEP_DT)
    st[:, 0] = identity_steps(b) if s.to_global else to_global_steps(pose, C * h * w)   # maps.py:2059-2060
    st[:, 1] = identity_steps(b) if target_to_global else from_global_steps(tpose, n_total)  # maps.py:2116-2117
    st = np.ascontiguousarray(st)
    n = h * w
    px = np.empty((b, C, n), np.float32)
    py = np.empty((b, C, n), np.float32)
    pz = np.empty((b, C, n), np.float32)
    woff, hoff = _vec(s.width_offset, b), _vec(s.height_offset, b)
    lib().dmo_fuse_points(_p(hm_c), ctypes.c_int64(C * n), ctypes.c_int64(n), b, C, h, w,
                          int(bool(s.flip_h)), ctypes.c_float(f32(s.map_res)), _p(woff), _p(hoff),
                          _p(st), _p(px), _p(py), _p(pz))
    pts_x.append(px); pts_y.append(py); pts_z.append(pz)
    masks.append(s.mask.reshape(b, C, n))
    if not is_height:
      vals.append(s.values.reshape(b, C, n))
  X = np.concatenate(pts_x, -1); Y = np.concatenate(pts_y, -1); Z = np.concatenate(pts_z, -1)
  Mk = np.concatenate(masks, -1)
  V = Y if is_height else np.concatenate(vals, -1)
  if Mk.sum() == 0:
    return None
  # _compute_new_shape_and_offsets, maps.py:2146-2179 (one bbox over everything valid)
  xv, zv = X[Mk].reshape(1, -1), Z[Mk].reshape(1, -1)
  xb, zb = map_quantize(xv, zv, 0., 0., target_res, None, False)
  min_x, max_x, min_z, max_z = int(xb.min()), int(xb.max()), int(zb.min()), int(zb.max())
  map_w = (max_x - min_x) + 2
  map_h = (max_z - min_z) + 2
  woff = np.float32(map_w / 2. - np.float32(np.float32(max_x + min_x) / np.float32(2.)))
  hoff = np.float32(map_h / 2. - np.float32(np.float32(max_z + min_z) / np.float32(2.)))
  b, C, NN = X.shape
  xb, zb = map_quantize(X.reshape(1, -1), Z.reshape(1, -1), woff, hoff, target_res, map_h, target_flip_h)
  coords = np.stack((zb.reshape(b, C, NN), xb.reshape(b, C, NN)), -1)
  fill = fill_value if fill_value is not None else (proj_fill_value if proj_fill_value is not None else NINF)
  canvas = np.zeros((b, C, map_h, map_w), np.float32)
  top, mask = scatter(V, coords, Mk, canvas, fill, reduction)
  if is_height:
    hgt = top
  else:
    hgt, _ = scatter(Y, coords, Mk, canvas, NINF, None)
  return dict(topdown=top, mask=mask, height=hgt, map_width=map_w, map_height=map_h,
              width_offset=woff, height_offset=hoff)


def fuse_inplace(world, source, canvas_shape, target_res, target_flip_h, fill_value=NINF, reduction=None):
  """Restatement of the opt-in fixed-canvas merge (dungeon_maps_b200.merge_into_canvas; NOT a reference
  function): the source map's valid cells become points as in fuse() above (maps.py:2039-2069), are
  quantised with the fixed offsets (Wc/2, Hc/2) (maps.py:944-1019) and scattered into the existing
  canvases with the reference's project(canvas=, canvas_masks=) rule (maps.py:1089-1173, utils.py:462-491):
  new = max(old, hits); mask |= changed.  `world` is None (fresh canvases) or the dict returned before."""
  Hc, Wc = canvas_shape
  s = source
  b, C, h, w = s.mask.shape
  is_height = s.values is None
  if world is None:
    world = dict(topdown=np.full((b, C, Hc, Wc), fill_value, np.float32), mask=np.zeros((b, C, Hc, Wc), bool),
                 height=None if is_height else np.full((b, C, Hc, Wc), NINF, np.float32))
  hm_c = np.ascontiguousarray(np.broadcast_to(s.height, (b, C, h, w)))
  pose = np.asarray(s.cam_pose, np.float32).reshape(-1, 3)
  if pose.shape[0] == 1 and b > 1:
    pose = np.repeat(pose, b, 0)
  n = h * w
  st = np.zeros((b, 2), dtype=STEP_DT)
  st[:, 0] = identity_steps(b) if s.to_global else to_global_steps(pose, C * n)
  st[:, 1] = identity_steps(b)
  st = np.ascontiguousarray(st)
  px = np.empty((b, C, n), np.float32); py = np.empty((b, C, n), np.float32); pz = np.empty((b, C, n), np.float32)
  woff, hoff = _vec(s.width_offset, b), _vec(s.height_offset, b)
  lib().dmo_fuse_points(_p(hm_c), ctypes.c_int64(C * n), ctypes.c_int64(n), b, C, h, w,
                        int(bool(s.flip_h)), ctypes.c_float(f32(s.map_res)), _p(woff), _p(hoff),
                        _p(st), _p(px), _p(py), _p(pz))
  xb, zb = map_quantize(px.reshape(1, -1), pz.reshape(1, -1), np.float32(Wc / 2.), np.float32(Hc / 2.), target_res,
                        Hc, target_flip_h)
  coords = np.stack((zb.reshape(b, C, n), xb.reshape(b, C, n)), -1)
  Mk = s.mask.reshape(b, C, n)
  V = py if is_height else s.values.reshape(b, C, n)
  top, changed = scatter(V, coords, Mk, world["topdown"], None, reduction)
  out = dict(topdown=top, mask=world["mask"] | changed, height=None)
  if not is_height:
    out["height"], _ = scatter(py, coords, Mk, world["height"], None, None)
  return out


def crop_nearest(image, center, crop_w, crop_h, fill_value):
  """image_sample(generate_crop_grid(...)), utils.py:571-652, on (b,c,h,w) float32."""
  image = np.ascontiguousarray(image, dtype=np.float32)
  b, c, h, w = image.shape
  center = np.ascontiguousarray(np.asarray(center, dtype=np.float32).reshape(-1, 2))
  if center.shape[0] == 1 and b > 1:
    center = np.ascontiguousarray(np.repeat(center, b, 0))
  out = np.empty((b, c, crop_h, crop_w), np.float32)
  lib().dmo_crop_nearest(_p(image), _p(center), b, c, h, w, crop_h, crop_w,
                         int(fill_value is not None),
                         ctypes.c_float(f32(0. if fill_value is None else fill_value)), _p(out))
  return out


def max_threads() -> int:
  return int(lib().dmo_max_threads())
