"""Drop-in for `dungeon_maps.maps` (reference: /root/reference/dungeon_maps/maps.py).

Same public functions and classes, same argument names / meaning / defaults;
the bodies call the fused sm_100a kernels through the C ABI
(include/dungeon_maps_b200.h) instead of chaining ~100 aten ops and
torch_scatter.  Host code here only validates and reshapes arguments, builds
the per-sample parameter blocks (_params.py) and allocates outputs.

There is no CPU path: inputs that are not on a CUDA device are moved to one
(`device=` or the current CUDA device); results live on that device.  Batch > 1
is supported and defined as "the reference applied to every sample"
(the reference itself raises for batch > 1, utils.py:311-316).
"""
import copy
import enum
import inspect
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _native as nat
from . import _params as prm
from . import utils
from .utils import CameraIntrinsics, Float3D, NINF, Reduction


@enum.unique
class CenterMode(str, enum.Enum):
  """maps.py:26-39; CenterMode(None) is CenterMode.none."""
  none = "none"
  origin = "origin"
  camera = "camera"

  @classmethod
  def _missing_(cls, value):
    if value is None:
      return cls.none


__all__ = [
  'CenterMode', 'get', 'orth_project', 'camera_affine_grid', 'compute_ego_flow', 'depth_map_to_point_cloud',
  'height_map_to_point_cloud', 'image_to_camera_space', 'camera_to_image_space', 'camera_to_local_space',
  'local_to_camera_space', 'local_to_global_space', 'global_to_local_space', 'map_quantize',
  'map_dequantize', 'project', 'compute_center_offsets', 'MapProjector', 'TopdownMap', 'crop_topdown_map',
  'fuse_topdown_maps', 'merge_into_canvas', 'MapBuilder', 'release_workspaces', 'Reduction', 'CameraIntrinsics', 'NINF', 'Float3D',
]


def get(*args: Any) -> Any:
  """First argument that is not None (the last one if all are)."""
  arg = None
  for arg in args:
    if arg is not None:
      break
  return arg


# ---- internal plumbing ----------------------------------------------------------------------

_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
  """Zero-initialised accumulation ring, one per (device, stream); the kernels leave it zeroed."""
  key = (dev.index, nat.stream_ptr(dev))
  ws = _workspaces.get(key)
  if ws is None or ws.numel() < nbytes:
    ws = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=dev)
    _workspaces[key] = ws
  return ws


def _alloc_canvas(shape: Tuple[int, ...], dtype: torch.dtype, dev: torch.device) -> torch.Tensor:
  """A fresh, contiguous tensor of `shape` for a map whose size is data dependent (the merged world map grows step by
  step, maps.py:2171-2178).  The storage is rounded up to a size class (a quarter of a power of two apart): torch's
  caching allocator can then hand the block a merge freed two steps ago to the next, slightly larger map instead of
  going to cudaMalloc — a device-wide synchronising call, twice per step — for every new size."""
  n = int(np.prod(shape))
  if n < (1 << 18):
    return torch.empty(shape, dtype=dtype, device=dev)
  return torch.empty((_canvas_cap(n),), dtype=dtype, device=dev)[:n].view(shape)


def _canvas_cap(n: int) -> int:
  """Size class of a canvas of n elements: the next of {5/8, 6/8, 7/8, 8/8} of the power of two above n."""
  top = 1 << (n - 1).bit_length()          # smallest power of two >= n
  return next(c for c in (top // 2 + k * (top // 8) for k in range(1, 5)) if c >= n)


def release_workspaces() -> None:
  """Frees the cached accumulation rings (one per device and stream that projected something) and the library's
  own scratch of the host-buffer entries."""
  _workspaces.clear()
  nat.lib().dm_release_scratch()


def _pick_device(device, *tensors) -> torch.device:
  if device is not None and not isinstance(device, bool):
    return nat.require_cuda(device)
  for t in tensors:
    if torch.is_tensor(t) and t.is_cuda:
      return t.device
  return nat.require_cuda(None)


def _image(x, dev: torch.device, dtype: torch.dtype) -> torch.Tensor:
  """2/3/4-D image → contiguous (b, c, h, w) on `dev`."""
  t = utils.to_tensor(x)
  t = utils.to_4D_image(t)
  return t.to(device=dev, dtype=dtype).contiguous()


def _points(x, dev: torch.device) -> Tuple[torch.Tensor, torch.Size]:
  """(..., 3) points → contiguous (b, n, 3) float32 on `dev` plus the original shape."""
  t = utils.to_tensor(x).to(device=dev, dtype=torch.float32)
  shape = t.shape
  if t.dim() < 2:
    t = t.view(-1, 3)
  b = t.shape[0]
  return t.reshape(b, -1, 3).contiguous(), shape


def _run_steps(points, dev, host_steps: List[torch.Tensor]) -> torch.Tensor:
  flat, shape = _points(points, dev)
  steps = torch.cat(host_steps, dim=1)
  return utils._apply_steps(flat, steps, len(host_steps)).reshape(shape)


def _f32(x) -> float:
  """Python float rounded to float32 (what torch does to a Python scalar in a float32 op)."""
  return float(np.float32(x))


# ======== Raw functional APIs (maps.py:121-1248) ==============================================

def orth_project(
  depth_map: torch.Tensor,
  value_map: Optional[torch.Tensor],
  valid_map: Optional[torch.Tensor],
  cam_pose: torch.Tensor,
  width_offset: torch.Tensor,
  height_offset: torch.Tensor,
  cam_pitch: torch.Tensor,
  cam_height: torch.Tensor,
  map_res: float,
  map_width: int,
  map_height: int,
  focal_x: float,
  focal_y: float,
  center_x: float,
  center_y: float,
  trunc_depth_min: Optional[float],
  trunc_depth_max: Optional[float],
  trunc_height_max: Optional[float],
  clip_border: Optional[int],
  to_global: bool,
  flip_h: bool = True,
  fill_value: Optional[float] = None,
  reduction: Optional[Reduction] = None,
  get_height_map: bool = False,
  device: Optional[torch.device] = None,
  _validate_args: bool = True,
  label_map: Optional[torch.Tensor] = None,
  num_classes: Optional[int] = None
) -> Union[Tuple[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
  """Orthographic projection of UNNORMALIZED depth maps (and optional per-pixel value maps)
  onto top-down maps: one fused kernel pass + one resolve pass (csrc/dm_project.cu) in place of
  maps.py:259-351.  Arguments, defaults and return values are those of the reference:

  Returns (topdown_map (b,C,mh,mw) f32, masks (b,C,mh,mw) bool[, height_map]); every value
  channel is reduced independently (max unless `reduction` says min); cells nothing landed on
  hold `fill_value` (0 when None) and are False in `masks`; `height_map` is the same tensor as
  `topdown_map` when `value_map` is None, otherwise a stride-0 expand of the (b,1,mh,mw)
  max-height map with -inf in empty cells.

  Addition to the reference's signature: `label_map` (b, 1, h, w) integer class ids with `num_classes` = C gives,
  bit for bit, the result of value_map = one_hot(label_map, C) as float32 planes (what the reference's object-map
  demo builds, demos/object_map/run.py:117-124) without ever materialising those planes — csrc/dm_labels.cu.
  """
  if label_map is not None:
    if value_map is not None:
      raise ValueError("pass either `value_map` or `label_map`, not both")
    if num_classes is None:
      raise ValueError("`label_map` needs `num_classes` (the C of one_hot(label_map, C))")
    if not 1 <= int(num_classes) <= 63:
      raise ValueError(f"num_classes must be in 1..63, got {num_classes}")
  if utils._reduction_code(reduction, fused=False) > 1:
    if label_map is not None:  # rarely used reductions: materialise the one-hot planes and take the composed path
      value_map = _one_hot_planes(label_map, int(num_classes), _pick_device(device, depth_map, label_map))
    return _orth_project_composed(
      depth_map, value_map, valid_map, cam_pose, width_offset, height_offset, cam_pitch, cam_height, map_res,
      map_width, map_height, focal_x, focal_y, center_x, center_y, trunc_depth_min, trunc_depth_max,
      trunc_height_max, clip_border, to_global, flip_h, fill_value, reduction, get_height_map, device)
  red = utils._reduction_code(reduction)
  dev = _pick_device(device, depth_map, value_map, label_map)
  depth = _image(depth_map, dev, torch.float32)
  b, dc, H, W = depth.shape
  values = None if value_map is None else _batch_of(_image(value_map, dev, torch.float32), b, "value_map")
  valid = None if valid_map is None else _batch_of(_image(valid_map, dev, torch.bool), b, "valid_map")
  labels = None if label_map is None else _batch_of(_label_ids(label_map, int(num_classes), dev), b, "label_map")
  for name, t in (("value_map", values), ("valid_map", valid), ("label_map", labels)):
    if t is not None and t.shape[-2:] != depth.shape[-2:]:
      raise RuntimeError(f"{name} {tuple(t.shape)} does not match depth_map {tuple(depth.shape)}")
  if labels is not None and (dc != 1 or labels.shape[1] != 1):
    raise RuntimeError("label_map needs single-channel depth and label maps: (b, 1, h, w)")
  # per-sample host parameters
  pose = prm.per_sample(cam_pose, b, (3,), "cam_pose")
  pitch = prm.per_sample(cam_pitch, b, (), "cam_pitch")
  camh = prm.per_sample(cam_height, b, (), "cam_height")
  woff = prm.per_sample(width_offset, b, (), "width_offset")
  hoff = prm.per_sample(height_offset, b, (), "height_offset")
  frames = b
  C = int(num_classes) if labels is not None else (0 if values is None else values.shape[1])
  if dc != 1:
    # one index set per depth channel (maps.py:298-318 keeps the channel dim): fold it into the batch
    if values is not None and values.shape[1] != dc:
      raise RuntimeError(f"value_map has {values.shape[1]} channels, depth_map has {dc}")
    if valid is not None and valid.shape[1] not in (1, dc):
      raise RuntimeError(f"valid_map has {valid.shape[1]} channels, depth_map has {dc}")
    frames = b * dc
    depth = depth.reshape(frames, 1, H, W)
    if values is not None:
      values = values.reshape(frames, 1, H, W)
      C = 1
    if valid is not None:
      valid = valid.expand(b, dc, H, W).reshape(frames, 1, H, W).contiguous()
    rep = lambda t: t.repeat_interleave(dc, dim=0)
    pose, pitch, camh, woff, hoff = rep(pose), rep(pitch), rep(camh), rep(woff), rep(hoff)
  elif valid is not None and valid.shape[1] != 1:
    raise RuntimeError(f"valid_map has {valid.shape[1]} channels, depth_map has 1")
  n_points = dc * H * W  # points the reference rotates per sample in one bmm
  samples, fast_steps = prm.proj_samples(pose, pitch, camh, woff, hoff, bool(to_global), n_points)
  samples_dev = prm.upload(samples, dev)

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
  want_height = bool(get_height_map) and C > 0
  cfg.want_height = want_height
  cfg.reduction = red
  cfg.fast_steps = fast_steps
  Cv = max(C, 1)
  topdown = torch.empty((frames, Cv, cfg.Mh, cfg.Mw), dtype=torch.float32, device=dev)
  masks = torch.empty((frames, Cv, cfg.Mh, cfg.Mw), dtype=torch.bool, device=dev)
  height = torch.empty((frames, 1, cfg.Mh, cfg.Mw), dtype=torch.float32, device=dev) if want_height else None
  lib = nat.lib()
  with torch.cuda.device(dev):
    if labels is not None:
      ws = _workspace(dev, lib.dm_orth_project_labels_workspace_bytes(cfg, frames))
      rc = lib.dm_orth_project_labels_f32(depth.data_ptr(), labels.data_ptr(), nat.ptr(valid), samples_dev.data_ptr(),
                                          cfg, frames, topdown.data_ptr(), masks.data_ptr(), nat.ptr(height),
                                          ws.data_ptr(), ws.numel(), nat.stream_ptr(dev))
      nat.check(rc, "dm_orth_project_labels_f32")
    else:
      ws = _workspace(dev, lib.dm_orth_project_workspace_bytes(cfg, frames))
      rc = lib.dm_orth_project_f32(depth.data_ptr(), nat.ptr(values), nat.ptr(valid), samples_dev.data_ptr(),
                                   cfg, frames, topdown.data_ptr(), masks.data_ptr(), nat.ptr(height),
                                   ws.data_ptr(), ws.numel(), nat.stream_ptr(dev))
      nat.check(rc, "dm_orth_project_f32")
  if dc != 1:
    topdown = topdown.reshape(b, dc, cfg.Mh, cfg.Mw)
    masks = masks.reshape(b, dc, cfg.Mh, cfg.Mw)
    if height is not None:
      height = height.reshape(b, dc, cfg.Mh, cfg.Mw)
  if not get_height_map:
    return topdown, masks
  if C == 0:
    return topdown, masks, topdown          # maps.py:333-334: the very same tensor
  return topdown, masks, torch.broadcast_to(height, topdown.shape)  # maps.py:349


def _batch_of(t: torch.Tensor, b: int, what: str) -> torch.Tensor:
  """A per-frame image tensor with batch 1 is shared by all b frames (the reference's ops broadcast it); the kernels
  index every plane by frame, so it is expanded here."""
  if t.shape[0] == b:
    return t
  if t.shape[0] == 1:
    return t.expand((b,) + tuple(t.shape[1:])).contiguous()
  raise RuntimeError(f"{what} has batch {t.shape[0]}, depth_map has batch {b}")


def _label_ids(label_map, num_classes: int, dev: torch.device) -> torch.Tensor:
  """Class ids as the kernels take them: contiguous (b, 1, h, w) uint8 on `dev`; ids outside [0, num_classes)
  become 255 ("no class": an all-zero one-hot row).  uint8 input is used as it is."""
  t = utils.to_4D_image(utils.to_tensor(label_map))
  if t.dtype is not torch.uint8:
    if t.dtype.is_floating_point or t.dtype is torch.bool:
      raise TypeError(f"label_map must hold integer class ids, got {t.dtype}")
    t = t.to(dev)
    t = torch.where((t < 0) | (t >= num_classes), torch.full_like(t, 255), t).to(torch.uint8)
  return t.to(dev).contiguous()


def _one_hot_planes(label_map, num_classes: int, dev: torch.device) -> torch.Tensor:
  ids = _label_ids(label_map, num_classes, dev)[:, 0].to(torch.int64)               # (b, h, w)
  planes = torch.arange(num_classes, device=dev).view(1, -1, 1, 1) == ids.unsqueeze(1)
  return planes.to(torch.float32)


def _orth_project_composed(depth_map, value_map, valid_map, cam_pose, width_offset, height_offset, cam_pitch,
                            cam_height, map_res, map_width, map_height, focal_x, focal_y, center_x, center_y,
                            trunc_depth_min, trunc_depth_max, trunc_height_max, clip_border, to_global, flip_h,
                            fill_value, reduction, get_height_map, device):
  """orth_project for the reductions the fused kernel does not cover (sum, mean, prod): the reference's own
  sequence of public steps (maps.py:259-351), each of them one kernel of this package — points are materialised
  here, the price of the rarely used reductions."""
  dev = _pick_device(device, depth_map, value_map)
  depth = _image(depth_map, dev, torch.float32)
  b, dc, H, W = depth.shape
  points, valid = depth_map_to_point_cloud(depth, valid_map, focal_x, focal_y, center_x, center_y, trunc_depth_min,
                                           trunc_depth_max, flip_h, device=dev)          # (b, c, h, w, 3)
  if clip_border is not None and clip_border > 0:                                        # maps.py:48-70
    k = int(clip_border)
    border = torch.zeros((H, W), dtype=torch.bool, device=dev)
    border[k:H - k, k:W - k] = True
    valid = valid & border
  points = camera_to_local_space(points, cam_pitch, cam_height, device=dev)
  if trunc_height_max is not None:                                                       # maps.py:286-288
    valid = valid & (points[..., 1] <= trunc_height_max)
  if to_global:
    points = local_to_global_space(points, cam_pose, device=dev)
  flat = points.reshape(b, dc, H * W, 3)
  flat_mask = valid.reshape(b, dc, H * W)
  x_bin, z_bin = map_quantize(flat[..., 0].reshape(b, -1), flat[..., 2].reshape(b, -1), width_offset, height_offset,
                              map_res, map_height, flip_h, device=dev)
  coords = torch.stack((z_bin, x_bin), dim=-1).reshape(b, dc, H * W, 2)
  heights = flat[..., 1]
  values = heights if value_map is None else _image(value_map, dev, torch.float32).reshape(b, -1, H * W)
  canvas = torch.zeros(values.shape[:-1] + (int(map_height), int(map_width)), dtype=torch.float32, device=dev)
  topdown, masks = project(coords, values, flat_mask, canvas, fill_value=fill_value, reduction=reduction, device=dev)
  if not get_height_map:
    return topdown, masks
  if value_map is None:
    return topdown, masks, topdown                                                       # maps.py:333-334
  hcanvas = torch.zeros((b, dc, int(map_height), int(map_width)), dtype=torch.float32, device=dev)
  height, _ = project(coords, heights, flat_mask, hcanvas, fill_value=NINF, reduction=Reduction.max, device=dev)
  return topdown, masks, torch.broadcast_to(height, topdown.shape)                       # maps.py:349


def camera_affine_grid(
  depth_map: torch.Tensor,
  trans_pose: torch.Tensor,
  cam_pitch: torch.Tensor,
  cam_height: torch.Tensor,
  focal_x: float,
  focal_y: float,
  center_x: float,
  center_y: float,
  flip_h: bool = True,
  device: Optional[torch.device] = None,
  _validate_args: bool = True,
  _emit_flow: bool = False
) -> torch.Tensor:
  """Where every pixel of `depth_map` (time t) lands in the image after the camera moved by
  `trans_pose` = [dx, dz, dyaw]: one fused element-wise kernel (csrc/dm_flow.cu) in place of
  maps.py:414-460.  Returns the (b, ..., h, w, 2) float32 grid of image (x, y)."""
  dev = _pick_device(device, depth_map)
  depth = _image(depth_map, dev, torch.float32)
  b, ch, H, W = depth.shape
  pose = prm.per_sample(trans_pose, b, (3,), "trans_pose")
  pitch = prm.per_sample(cam_pitch, b, (), "cam_pitch")
  camh = prm.per_sample(cam_height, b, (), "cam_height")
  n_points = ch * H * W
  samples_dev = prm.upload(prm.flow_samples(pose, pitch, camh, n_points), dev)
  cfg = nat.DmFlowCfg()
  cfg.H, cfg.W, cfg.channels = H, W, ch
  cfg.fx, cfg.fy, cfg.cx, cfg.cy = focal_x, focal_y, center_x, center_y
  cfg.flip_h = bool(flip_h)
  cfg.emit_flow = bool(_emit_flow)
  grid = torch.empty((b, ch, H, W, 2), dtype=torch.float32, device=dev)
  with torch.cuda.device(dev):
    rc = nat.lib().dm_affine_grid_f32(depth.data_ptr(), samples_dev.data_ptr(), cfg, b, grid.data_ptr(),
                                      nat.stream_ptr(dev))
  nat.check(rc, "dm_affine_grid_f32")
  return grid


def compute_ego_flow(proj: "MapProjector", depth_map: torch.Tensor, trans_pose: torch.Tensor) -> torch.Tensor:
  """Egocentric motion flow of the reference's demo helper (demos/ego_flow/run.py:75-90):
  (x - grid_x, -(y - grid_y)) in pixels, fused into the same kernel.  depth_map (c, h, w) or
  (b, c, h, w); returns the flow of the first sample / channel, (h, w, 2), like the demo."""
  flow = camera_affine_grid(
    depth_map=depth_map, trans_pose=trans_pose, cam_pitch=proj.cam_pitch, cam_height=proj.cam_height,
    focal_x=proj.cam_params.fx, focal_y=proj.cam_params.fy, center_x=proj.cam_params.cx,
    center_y=proj.cam_params.cy, flip_h=proj.flip_h, device=proj.device, _emit_flow=True)
  return flow[0, 0]


def depth_map_to_point_cloud(
  depth_map: torch.Tensor,
  valid_map: Optional[torch.Tensor],
  focal_x: float,
  focal_y: float,
  center_x: float,
  center_y: float,
  trunc_depth_min: Optional[float],
  trunc_depth_max: Optional[float],
  flip_h: bool = True,
  device: Optional[torch.device] = None,
  _validate_args: bool = True
) -> Tuple[torch.Tensor, torch.Tensor]:
  """Camera-space points (b, c, h, w, 3) and the valid area (b, c, h, w) of a depth map
  (maps.py:462-545); X right, Y up, Z forward."""
  dev = _pick_device(device, depth_map)
  depth = _image(depth_map, dev, torch.float32)
  b, c, H, W = depth.shape
  valid = None
  if valid_map is not None:
    valid = _image(valid_map, dev, torch.bool).expand(b, c, H, W).contiguous()
  pts = torch.empty((b, c, H, W, 3), dtype=torch.float32, device=dev)
  ok = torch.empty((b, c, H, W), dtype=torch.bool, device=dev)
  with torch.cuda.device(dev):
    rc = nat.lib().dm_depth_to_points_f32(
      depth.data_ptr(), nat.ptr(valid), b * c, H, W, focal_x, focal_y, center_x, center_y, int(bool(flip_h)),
      int(trunc_depth_min is not None), trunc_depth_min or 0., int(trunc_depth_max is not None),
      trunc_depth_max or 0., pts.data_ptr(), ok.data_ptr(), nat.stream_ptr(dev))
  nat.check(rc, "dm_depth_to_points_f32")
  return pts, ok


def height_map_to_point_cloud(
  height_map: torch.Tensor,
  width_offset: torch.Tensor,
  height_offset: torch.Tensor,
  map_res: float,
  map_height: int,
  flip_h: bool = True,
  device: Optional[torch.device] = None,
  _validate_args: bool = True
) -> torch.Tensor:
  """Every cell of a height map as a 3-D point (b, c, h, w, 3) (maps.py:547-612)."""
  dev = _pick_device(device, height_map)
  hm = _image(height_map, dev, torch.float32)
  xb, zb = utils.generate_image_coords(hm.shape, dtype=torch.float32, device=dev)
  x, z = map_dequantize(xb, zb, width_offset, height_offset, map_res, map_height, flip_h, device=dev)
  return torch.stack((x, hm, z), dim=-1)


def _image_camera(points, fx, fy, cx, cy, flip_h, height, to_image, device):
  dev = _pick_device(device, points)
  t = utils.to_tensor(points).to(device=dev, dtype=torch.float32)
  if flip_h and height is None:
    if t.dim() < 3:
      raise RuntimeError("The rank of `points` must be at least 3D (..., h, w, 3) "
                         "or `height` should be provided if `flip_h` is enabled.")
    height = t.shape[-3]
  flat = t.reshape(-1, 3).contiguous()
  out = torch.empty_like(flat)
  with torch.cuda.device(dev):
    rc = nat.lib().dm_image_camera_f32(flat.data_ptr(), flat.shape[0], fx, fy, cx, cy, int(bool(flip_h)),
                                       int(height or 0), int(to_image), out.data_ptr(), nat.stream_ptr(dev))
  nat.check(rc, "dm_image_camera_f32")
  return out.reshape(t.shape)


def image_to_camera_space(points: torch.Tensor, focal_x: float, focal_y: float, center_x: float,
                          center_y: float, flip_h: bool = True, height: Optional[int] = None,
                          device: Optional[torch.device] = None, _validate_args: bool = True) -> torch.Tensor:
  """(pixel x, pixel y, depth) → camera space (maps.py:616-682)."""
  return _image_camera(points, focal_x, focal_y, center_x, center_y, flip_h, height, 0, device)


def camera_to_image_space(points: torch.Tensor, focal_x: float, focal_y: float, center_x: float,
                          center_y: float, flip_h: bool = True, height: Optional[int] = None,
                          device: Optional[torch.device] = None, _validate_args: bool = True) -> torch.Tensor:
  """Camera space → (pixel x, pixel y, depth) (maps.py:684-751)."""
  return _image_camera(points, focal_x, focal_y, center_x, center_y, flip_h, height, 1, device)


def _n_points(points) -> Tuple[int, int]:
  t = utils.to_tensor(points)
  if t.dim() < 2:
    return 1, max(t.numel() // 3, 1)
  b = t.shape[0]
  return b, max(t.numel() // (3 * max(b, 1)), 1)


def _per_sample_steps(points, build, *params):
  """Step blocks for a transform whose per-sample parameters (pose, pitch, height) may carry a batch of their own:
  like the reference's tensor ops, a single point set is broadcast over b parameter sets (TopdownMap.get_camera of
  a batched map, maps.py:1824-1839) and a single parameter set over b point sets.  Returns (points, steps)."""
  bp, n = _n_points(points)
  tails = [(3,) if isinstance(p, tuple) else () for p in params]
  vals = [prm.host_f32(p[0] if isinstance(p, tuple) else p, tail) for p, tail in zip(params, tails)]
  b = max([bp] + [v.shape[0] for v in vals])
  if bp != b:
    t = utils.to_tensor(points)
    if bp != 1:
      raise ValueError(f"points have batch {bp}, the transform parameters {b}")
    if t.dim() < 2:
      t = t.view(1, -1, 3)
    points = t.expand((b,) + tuple(t.shape[1:]))
  args = [prm.per_sample(v, b, tail) for v, tail in zip(vals, tails)]
  return points, build(*args, n)


def camera_to_local_space(points: torch.Tensor, cam_pitch: torch.Tensor, cam_height: torch.Tensor,
                          device: Optional[torch.device] = None, _validate_args: bool = True) -> torch.Tensor:
  """Rotate by the camera pitch about x, lift by the camera height (maps.py:753-800)."""
  points, st = _per_sample_steps(points, prm.camera_to_local, cam_pitch, cam_height)
  return _run_steps(points, _pick_device(device, points), [st])


def local_to_camera_space(points: torch.Tensor, cam_pitch: torch.Tensor, cam_height: torch.Tensor,
                          device: Optional[torch.device] = None, _validate_args: bool = True) -> torch.Tensor:
  """Inverse of camera_to_local_space (maps.py:802-848)."""
  points, st = _per_sample_steps(points, prm.local_to_camera, cam_pitch, cam_height)
  return _run_steps(points, _pick_device(device, points), [st])


def local_to_global_space(points: torch.Tensor, cam_pose: torch.Tensor,
                          device: Optional[torch.device] = None, _validate_args: bool = True) -> torch.Tensor:
  """Rotate by yaw about y, translate by (x, 0, z) (maps.py:850-895)."""
  points, st = _per_sample_steps(points, prm.local_to_global, (cam_pose,))
  return _run_steps(points, _pick_device(device, points), [st])


def global_to_local_space(points: torch.Tensor, cam_pose: torch.Tensor,
                          device: Optional[torch.device] = None, _validate_args: bool = True) -> torch.Tensor:
  """Inverse of local_to_global_space (maps.py:897-942)."""
  points, st = _per_sample_steps(points, prm.global_to_local, (cam_pose,))
  return _run_steps(points, _pick_device(device, points), [st])


def _coords_pair(x_coords, z_coords, dev):
  x = utils.to_tensor(x_coords).to(device=dev, dtype=torch.float32)
  z = utils.to_tensor(z_coords).to(device=dev, dtype=torch.float32)
  x, z = torch.broadcast_tensors(x, z)
  shape = x.shape
  if x.dim() < 2:
    x, z = x.reshape(1, -1), z.reshape(1, -1)
  b = x.shape[0]
  return x.reshape(b, -1).contiguous(), z.reshape(b, -1).contiguous(), shape, b


def map_quantize(x_coords: torch.Tensor, z_coords: torch.Tensor, width_offset: torch.Tensor,
                 height_offset: torch.Tensor, map_res: float, map_height: Optional[int] = None,
                 flip_h: bool = True, device: Optional[torch.device] = None,
                 _validate_args: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
  """World x, z → integer map bins, rounding half up (maps.py:944-1019)."""
  dev = _pick_device(device, x_coords, z_coords)
  x, z, shape, b = _coords_pair(x_coords, z_coords, dev)
  if flip_h:
    assert map_height is not None
  woff = prm.upload(prm.per_sample(width_offset, b).contiguous(), dev)
  hoff = prm.upload(prm.per_sample(height_offset, b).contiguous(), dev)
  xb = torch.empty(x.shape, dtype=torch.int64, device=dev)
  zb = torch.empty(x.shape, dtype=torch.int64, device=dev)
  with torch.cuda.device(dev):
    rc = nat.lib().dm_map_quantize_f32(x.data_ptr(), z.data_ptr(), woff.data_ptr(), hoff.data_ptr(), b,
                                       x.shape[1], map_res, int(map_height or 0), int(bool(flip_h)),
                                       xb.data_ptr(), zb.data_ptr(), nat.stream_ptr(dev))
  nat.check(rc, "dm_map_quantize_f32")
  out_shape = shape if len(shape) >= 2 else (1, xb.numel())
  return xb.reshape(out_shape), zb.reshape(out_shape)


def map_dequantize(x_coords: torch.Tensor, z_coords: torch.Tensor, width_offset: torch.Tensor,
                   height_offset: torch.Tensor, map_res: float, map_height: Optional[int] = None,
                   flip_h: bool = True, device: Optional[torch.device] = None,
                   _validate_args: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
  """Inverse of map_quantize (maps.py:1021-1087)."""
  dev = _pick_device(device, x_coords, z_coords)
  xb, zb, shape, b = _coords_pair(x_coords, z_coords, dev)
  if flip_h:
    assert map_height is not None
  woff = prm.upload(prm.per_sample(width_offset, b).contiguous(), dev)
  hoff = prm.upload(prm.per_sample(height_offset, b).contiguous(), dev)
  x = torch.empty_like(xb)
  z = torch.empty_like(zb)
  with torch.cuda.device(dev):
    rc = nat.lib().dm_map_dequantize_f32(xb.data_ptr(), zb.data_ptr(), woff.data_ptr(), hoff.data_ptr(), b,
                                         xb.shape[1], map_res, int(map_height or 0), int(bool(flip_h)),
                                         x.data_ptr(), z.data_ptr(), nat.stream_ptr(dev))
  nat.check(rc, "dm_map_dequantize_f32")
  out_shape = shape if len(shape) >= 2 else (1, x.numel())
  return x.reshape(out_shape), z.reshape(out_shape)


def project(coords: torch.Tensor, values: torch.Tensor, masks: torch.Tensor, canvas: torch.Tensor,
            canvas_masks: Optional[torch.Tensor] = None, fill_value: Optional[float] = None,
            reduction: Optional[Reduction] = None, device: Optional[torch.device] = None,
            _validate_args: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
  """Scatter `values` (b, ..., n) onto `canvas` (b, ..., mh, mw) at `coords` (b, ..., n, 2) =
  [row, col]; out-of-range or masked points are dropped (maps.py:1089-1173)."""
  dev = _pick_device(device, coords, values, canvas)
  coords = utils.to_tensor(coords).to(device=dev, dtype=torch.int64)
  if coords.dim() < 3:
    coords = coords.view(1, -1, 2)
  maps_, changed = utils.scatter_tensor(canvas=canvas, indices=coords, values=values, masks=masks,
                                        fill_value=fill_value, reduction=reduction)
  if canvas_masks is not None:
    cm = utils.to_tensor(canvas_masks).to(device=dev, dtype=torch.bool)
    changed = torch.logical_or(torch.broadcast_to(cm, changed.shape), changed)
  return maps_, changed


def compute_center_offsets(cam_pose: torch.Tensor, width_offset: torch.Tensor, height_offset: torch.Tensor,
                           map_res: float, map_width: float, map_height: int, to_global: bool,
                           center_mode: CenterMode = CenterMode.none, device: Optional[torch.device] = None,
                           _validate_args: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
  """Map offsets that put the origin / the camera at the map centre (maps.py:1175-1248).
  A handful of floats per sample: host arithmetic (float32, reference op order); the results
  are host tensors that feed the kernels' parameter blocks."""
  center_mode = CenterMode(center_mode)
  pose = prm.host_f32(np.zeros(3, np.float32) if cam_pose is None else cam_pose)
  woff = prm.host_f32(0. if width_offset is None else width_offset)
  hoff = prm.host_f32(0. if height_offset is None else height_offset)
  if center_mode is CenterMode.none:
    return woff + 0., hoff + 0.
  pose = pose.reshape(-1, 3)
  center_x = torch.zeros(pose.shape[0])
  center_z = torch.zeros(pose.shape[0])
  if center_mode is CenterMode.camera and to_global:
    # local_to_global_space of the origin: rot(0) + (x, 0, z)   (maps.py:1228-1233)
    center_x = center_x + pose[:, 0]
    center_z = center_z + pose[:, 1]
  # map_quantize(width_offset=0., height_offset=0., flip_h=False)   (maps.py:1235-1243)
  res = torch.tensor(map_res, dtype=torch.float32)
  qx = torch.floor(center_x / res + 0. + 0.5).to(torch.int64).view(1, -1)
  qz = torch.floor(center_z / res + 0. + 0.5).to(torch.int64).view(1, -1)
  return woff + (map_width / 2. - qx), hoff + (map_height / 2. - qz)


# ======== MapProjector (maps.py:1252-1749) ======================================================

# name of a functional-API parameter → where MapProjector finds its default
_INTRINSIC_DEFAULTS = {"focal_x": "fx", "focal_y": "fy", "center_x": "cx", "center_y": "cy"}
_ATTR_DEFAULTS = (
  "cam_pose", "width_offset", "height_offset", "cam_pitch", "cam_height", "map_res", "map_width",
  "map_height", "trunc_depth_min", "trunc_depth_max", "trunc_height_max", "clip_border", "to_global",
  "flip_h", "fill_value", "reduction", "device", "height",
)
_OPTIONAL_DATA = ("value_map", "valid_map", "label_map", "num_classes")
_CTOR_ARGS = (
  "width", "height", "hfov", "vfov", "cam_pose", "width_offset", "height_offset", "cam_pitch", "cam_height",
  "map_res", "map_width", "map_height", "trunc_depth_min", "trunc_depth_max", "trunc_height_max",
  "clip_border", "to_global", "flip_h", "fill_value", "reduction", "device",
)
_CTOR_ARG_SET = frozenset(_CTOR_ARGS)


def _with_projector_defaults(fn):
  """Method wrapper: every argument left at None falls back to the projector's stored default
  (the reference spells this out per method as get(arg, self.arg), maps.py:1406-1749)."""
  params = list(inspect.signature(fn).parameters.values())
  names = [p.name for p in params]
  own_defaults = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
  keep_own = {"get_height_map", "_validate_args", "center_mode", "canvas_masks", "_emit_flow"}

  def method(self, *args, **kwargs):
    if len(args) > len(names):
      raise TypeError(f"{fn.__name__}() takes at most {len(names)} positional arguments")
    call = dict(zip(names, args))
    for k, v in kwargs.items():
      if k not in names:
        raise TypeError(f"{fn.__name__}() got an unexpected keyword argument '{k}'")
      if k in call:
        raise TypeError(f"{fn.__name__}() got multiple values for argument '{k}'")
      call[k] = v
    for name in names:
      if name in keep_own:
        call.setdefault(name, own_defaults[name])
      elif call.get(name) is None:
        if name in _INTRINSIC_DEFAULTS:
          call[name] = getattr(self.cam_params, _INTRINSIC_DEFAULTS[name])
        elif name in _ATTR_DEFAULTS:
          call[name] = getattr(self, name)
        elif name in _OPTIONAL_DATA:
          call[name] = None
        elif name not in call:
          raise TypeError(f"{fn.__name__}() missing required argument: '{name}'")
    return fn(**call)

  method.__name__ = fn.__name__
  method.__doc__ = fn.__doc__
  return method


class MapProjector():
  """Stores the camera / map configuration once so the functional APIs can be called with only
  the per-frame data; per-call keyword arguments override the stored values.  Constructor
  arguments are those of the reference (maps.py:1253-1347): width, height, hfov, vfov, cam_pose
  [x, z, yaw], width_offset, height_offset, cam_pitch, cam_height, map_res, map_width, map_height,
  trunc_depth_min/max, trunc_height_max, clip_border, to_global, flip_h, fill_value (default
  -inf), reduction, device."""

  def __init__(
    self,
    width: int,
    height: int,
    hfov: float,
    vfov: Optional[float] = None,
    cam_pose: Optional[Float3D] = None,
    width_offset: Optional[float] = None,
    height_offset: Optional[float] = None,
    cam_pitch: Optional[float] = None,
    cam_height: Optional[float] = None,
    map_res: Optional[float] = None,
    map_width: Optional[int] = None,
    map_height: Optional[int] = None,
    trunc_depth_min: Optional[float] = None,
    trunc_depth_max: Optional[float] = None,
    trunc_height_max: Optional[float] = None,
    clip_border: Optional[int] = None,
    to_global: bool = False,
    flip_h: bool = True,
    fill_value: Optional[float] = NINF,
    reduction: Optional[Reduction] = None,
    device: Optional[torch.device] = None
  ):
    given = locals()
    for name in _CTOR_ARGS:
      setattr(self, name, given[name])
    self.cam_params: CameraIntrinsics = utils.get_camera_intrinsics(
      width=self.width, height=self.height, hfov=self.hfov, vfov=self.vfov)

  def clone(self, **overrides) -> "MapProjector":
    """Shallow copy with some constructor arguments replaced (None keeps the stored value),
    maps.py:1349-1404."""
    other = copy.copy(self)  # no constructor run: clones happen several times per MapBuilder step
    intrinsics_changed = False
    for name, value in overrides.items():
      if name not in _CTOR_ARG_SET:
        raise TypeError(f"clone() got unexpected keyword arguments {sorted(set(overrides) - _CTOR_ARG_SET)}")
      if value is not None:
        setattr(other, name, value)
        intrinsics_changed = intrinsics_changed or name in ("width", "height", "hfov", "vfov")
    if intrinsics_changed:
      other.cam_params = utils.get_camera_intrinsics(width=other.width, height=other.height, hfov=other.hfov,
                                                     vfov=other.vfov)
    return other

  orth_project = _with_projector_defaults(orth_project)
  camera_affine_grid = _with_projector_defaults(camera_affine_grid)
  depth_map_to_point_cloud = _with_projector_defaults(depth_map_to_point_cloud)
  height_map_to_point_cloud = _with_projector_defaults(height_map_to_point_cloud)
  image_to_camera_space = _with_projector_defaults(image_to_camera_space)
  camera_to_image_space = _with_projector_defaults(camera_to_image_space)
  camera_to_local_space = _with_projector_defaults(camera_to_local_space)
  local_to_camera_space = _with_projector_defaults(local_to_camera_space)
  local_to_global_space = _with_projector_defaults(local_to_global_space)
  global_to_local_space = _with_projector_defaults(global_to_local_space)
  map_quantize = _with_projector_defaults(map_quantize)
  map_dequantize = _with_projector_defaults(map_dequantize)
  project = _with_projector_defaults(project)
  compute_center_offsets = _with_projector_defaults(compute_center_offsets)


# ======== TopdownMap (maps.py:1753-1955) ========================================================

class TopdownMap():
  """A top-down map with its mask, height map and the projector that produced it."""

  def __init__(self, topdown_map: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
               height_map: Optional[torch.Tensor] = None, map_projector: Optional[MapProjector] = None,
               is_height_map: Optional[bool] = None):
    self._proj = map_projector
    self._topdown_map = topdown_map
    self._mask = mask
    self._height_map = height_map
    if is_height_map is None:  # same object ⇒ the map is its own height map (maps.py:1781-1783)
      is_height_map = (topdown_map is not None) and (topdown_map is height_map)
    self._is_height_map = is_height_map

  @property
  def is_empty(self) -> bool:
    return self._topdown_map is None

  @property
  def is_height_map(self) -> bool:
    return self._is_height_map

  @property
  def map(self) -> torch.Tensor:
    return self._topdown_map

  @property
  def topdown_map(self) -> torch.Tensor:
    return self._topdown_map

  @property
  def height_map(self) -> torch.Tensor:
    return self._topdown_map if self._is_height_map else self._height_map

  @property
  def mask(self) -> torch.Tensor:
    return self._mask

  @property
  def proj(self) -> MapProjector:
    return self._proj

  def get_camera(self) -> torch.Tensor:
    """Map coordinates (b, 2) int64 of the camera, i.e. of the local origin (maps.py:1824-1839)."""
    return self.get_coords(torch.zeros((3,), dtype=torch.float32), is_global=False).squeeze(dim=-2)

  def get_origin(self) -> torch.Tensor:
    """Map coordinates (b, 2) int64 of the global origin (maps.py:1841-1856)."""
    return self.get_coords(torch.zeros((3,), dtype=torch.float32), is_global=True).squeeze(dim=-2)

  def get_coords(self, points: torch.Tensor, is_global: bool = True) -> torch.Tensor:
    """Map coordinates (b, n, 2) int64 [x_bin, z_bin] of 3-D points (maps.py:1858-1897)."""
    points = utils.to_tensor(points)
    if points.dim() < 3:
      points = points.view(1, -1, 3)
    if self.proj.to_global and not is_global:
      points = self.proj.local_to_global_space(points=points)
    elif (not self.proj.to_global) and is_global:
      points = self.proj.global_to_local_space(points=points)
    pos_x, pos_z = self.proj.map_quantize(x_coords=points[..., 0], z_coords=points[..., 2])
    return torch.stack((pos_x, pos_z), dim=-1)

  def get_points(self, coords: torch.Tensor) -> torch.Tensor:
    """World (x, z) (b, n, 2) float32 of map coordinates (maps.py:1899-1921)."""
    coords = utils.to_tensor(coords)
    if coords.dim() < 3:
      coords = coords.view(1, -1, 2)
    pos_x, pos_z = self.proj.map_dequantize(x_coords=coords[..., 0], z_coords=coords[..., 1])
    return torch.stack((pos_x, pos_z), dim=-1)

  def select(self, center: torch.Tensor, crop_width: int, crop_height: int,
             fill_value: Optional[float] = None) -> "TopdownMap":
    """Crop (or pad) a crop_width × crop_height window around `center` (maps.py:1923-1949)."""
    return crop_topdown_map(self, center=center, crop_width=crop_width, crop_height=crop_height,
                            fill_value=fill_value, _validate_args=True)

  def merge(self, *sources: List["TopdownMap"]) -> "TopdownMap":
    raise NotImplementedError


# ======== TopdownMap functional API ===============================================================

def _crop(image: torch.Tensor, center_dev: torch.Tensor, crop_h: int, crop_w: int,
          fill_value: Optional[float]) -> torch.Tensor:
  b, c, h, w = image.shape
  dev = image.device
  lib = nat.lib()
  with torch.cuda.device(dev):
    if image.dtype == torch.bool:
      out = torch.empty((b, c, crop_h, crop_w), dtype=torch.bool, device=dev)
      rc = lib.dm_crop_nearest_u8(image.data_ptr(), center_dev.data_ptr(), b, c, h, w, crop_h, crop_w,
                                  out.data_ptr(), nat.stream_ptr(dev))
    else:
      out = torch.empty((b, c, crop_h, crop_w), dtype=torch.float32, device=dev)
      rc = lib.dm_crop_nearest_f32(image.data_ptr(), center_dev.data_ptr(), b, c, h, w, crop_h, crop_w,
                                   int(fill_value is not None), 0. if fill_value is None else fill_value,
                                   out.data_ptr(), nat.stream_ptr(dev))
  nat.check(rc, "dm_crop_nearest")
  return out


def crop_topdown_map(source: TopdownMap, center: torch.Tensor, crop_width: int, crop_height: int,
                     fill_value: Optional[float] = None, mode: str = 'nearest',
                     _validate_args: bool = True) -> TopdownMap:
  """Nearest-neighbour crop of map, mask and height map around `center` (b, 2) = [x, y] in map
  pixels, exactly as the reference's pad + grid_sample formulation resamples it
  (maps.py:1959-2037); one fused kernel per tensor (csrc/dm_points.cu crop_kernel).  Unlike the
  reference, the caller's `center` tensor is not modified (maps.py:2021-2022 does, by aliasing)."""
  if mode != 'nearest':
    raise NotImplementedError("only mode='nearest' is implemented by the B200 kernels")
  proj = source.proj
  hm = source.height_map
  dev = hm.device
  center_host = prm.host_f32(center, (2,))
  b = hm.shape[0]
  center_dev = prm.upload(prm.per_sample(center_host, b, (2,), "center").contiguous(), dev)
  base = hm[:, :1] if (hm.dim() == 4 and hm.stride(1) == 0 and hm.shape[1] > 1) else hm
  height_map = _crop(base.contiguous(), center_dev, crop_height, crop_width, NINF)
  if base is not hm:
    height_map = height_map.expand(b, hm.shape[1], crop_height, crop_width)
  mask = _crop(source.mask.to(torch.bool).contiguous(), center_dev, crop_height, crop_width, False)
  topdown_map = height_map
  if not source.is_height_map:
    topdown_map = _crop(source.topdown_map.contiguous(), center_dev, crop_height, crop_width,
                        get(fill_value, proj.fill_value))
  # new offsets (maps.py:2020-2024), float32 host arithmetic
  cx, cy = center_host[:, 0].clone(), center_host[:, 1].clone()
  if proj.flip_h:
    cy = (proj.map_height - 1) - cy
  width_offset = prm.host_f32(proj.width_offset) + crop_width / 2 - cx
  height_offset = prm.host_f32(proj.height_offset) + crop_height / 2 - cy
  new_proj = proj.clone(width_offset=width_offset, height_offset=height_offset, map_width=crop_width,
                        map_height=crop_height)
  return TopdownMap(topdown_map=topdown_map, mask=mask, height_map=height_map,
                    is_height_map=source.is_height_map, map_projector=new_proj)


def _fuse_source(m: TopdownMap, target: MapProjector, b: int, C: int, n_total: int, dev: torch.device,
                 keep: list) -> nat.DmFuseSource:
  hm = m.height_map.to(device=dev, dtype=torch.float32)
  if hm.dim() != 4:
    hm = utils.to_4D_image(hm)
  h, w = hm.shape[-2:]
  if hm.stride(-1) != 1 or hm.stride(-2) != w:
    hm = hm.contiguous()
  hm = hm.expand(b, C, h, w)
  mask = m.mask.to(device=dev, dtype=torch.bool).expand(b, C, h, w).contiguous()
  values = None
  if not m.is_height_map:
    values = m.topdown_map.to(device=dev, dtype=torch.float32).expand(b, C, h, w).contiguous()
  pose = prm.per_sample(get(m.proj.cam_pose, [0., 0., 0.]), b, (3,), "cam_pose")
  tpose = prm.per_sample(get(target.cam_pose, [0., 0., 0.]), b, (3,), "cam_pose")
  # maps.py:2059-2060: a local map goes to global space with its own pose;
  # maps.py:2116-2117: everything goes to the target's local space if the target is local
  s0 = prm.identity(b) if m.proj.to_global else prm.local_to_global(pose, C * h * w)
  s1 = prm.identity(b) if target.to_global else prm.global_to_local(tpose, n_total)
  # one parameter block per source: [steps (b, 2, 16 words) | width offsets (b,) | height offsets (b,)], one upload
  params = torch.cat((torch.cat((s0, s1), dim=1).reshape(-1),
                      prm.per_sample(get(m.proj.width_offset, 0.), b).reshape(-1),
                      prm.per_sample(get(m.proj.height_offset, 0.), b).reshape(-1)))
  params_dev = prm.upload(params, dev)
  n_step_words = b * 2 * prm.STEP_WORDS
  keep.extend((hm, mask, values, params_dev))
  src = nat.DmFuseSource()
  src.height, src.values, src.mask = hm.data_ptr(), nat.ptr(values), mask.data_ptr()
  src.height_bstride, src.height_cstride = hm.stride(0), hm.stride(1)
  src.h, src.w = h, w
  src.flip_h = bool(m.proj.flip_h)
  src.map_res = m.proj.map_res
  base = params_dev.data_ptr()
  src.steps, src.width_offset, src.height_offset = base, base + 4 * n_step_words, base + 4 * (n_step_words + b)
  # a map this module wrote carries, per plane, the rectangle that holds its valid cells: the passes scan that instead
  # of the whole plane (a grown world map is mostly empty canvas)
  tb = getattr(m, "_tracked_box", None)
  if (tb is not None and tb.plane_box is not None and m.mask is tb.mask and tb.mask._version == tb.mask_version
      and tb.plane_box.device == dev and tuple(m.mask.shape) == (b, C, h, w)):
    src.plane_box = tb.plane_box.data_ptr()
    keep.append(tb.plane_box)
  return src


def fuse_topdown_maps(*maps: List[TopdownMap], map_projector: Optional[MapProjector] = None,
                      fill_value: Optional[float] = None, reduction: Optional[Reduction] = None) -> TopdownMap:
  """Re-project several top-down maps into one freshly sized map in `map_projector`'s frame
  (maps.py:2181-2287): every valid cell becomes a point again, one bounding box over all of them
  sizes the canvas (host sync), and the points are scatter-maxed into it.  Two fused passes over
  the source cells (csrc/dm_fuse.cu), no point cloud is materialised."""
  if len(maps) == 0:
    return TopdownMap(map_projector=map_projector)
  proj = map_projector if map_projector is not None else maps[0].proj
  live = [m for m in maps if not m.is_empty]
  if not live:
    return TopdownMap(map_projector=proj)
  kinds = {bool(m.is_height_map) for m in live}
  assert len(kinds) == 1, "All maps must be the same type of maps (all height maps or all value maps)."
  is_height_map = kinds.pop()
  # MapProjector.project: get(reduction, self.reduction), maps.py:1720.  All five reductions: max / min through the
  # atomic scatter, sum / mean / prod through the ordered fold (csrc/dm_ordered.cu: the reference's point order)
  red = utils._reduction_code(get(reduction, proj.reduction), fused=False)
  dev = _pick_device(proj.device, *[m.mask for m in live], *[m.height_map for m in live])
  shapes = [utils.to_4D_image(m.mask).shape for m in live]
  b = max(s[0] for s in shapes)
  C = max(s[1] for s in shapes)
  n_total = sum(C * s[2] * s[3] for s in shapes)
  keep: list = []
  # A map that an earlier call of this function produced carries the bounding box pass 1 would find for it
  # (_TrackedBox); it then only seeds the box and pass 1 scans the other maps — for a MapBuilder that is the local
  # map instead of the whole world map, ahead of the host sync below.
  seed = next((m for m in live if _tracked_box_valid(m, proj, dev)), None)
  scan = [m for m in live if m is not seed]
  all_sources = [_fuse_source(m, proj, b, C, n_total, dev, keep) for m in live]
  sources = (nat.DmFuseSource * len(live))(*all_sources)
  scan_sources = (nat.DmFuseSource * max(len(scan), 1))(*[s for m, s in zip(live, all_sources) if m is not seed])
  lib = nat.lib()
  bbox = torch.empty((5,), dtype=torch.int64, device=dev)
  with torch.cuda.device(dev):
    rc = lib.dm_fuse_bbox_seeded_i64(scan_sources, len(scan), b, C, proj.map_res,
                                     None if seed is None else seed._tracked_box.box.data_ptr(), bbox.data_ptr(),
                                     nat.stream_ptr(dev))
  nat.check(rc, "dm_fuse_bbox_seeded_i64")
  min_x, max_x, min_z, max_z, n_valid = (int(v) for v in bbox.cpu())  # the reference's .item() sync
  if n_valid == 0:  # maps.py:2217-2225
    last = maps[-1]
    return TopdownMap(topdown_map=last.topdown_map, mask=last.mask, height_map=last.height_map, map_projector=proj)
  # maps.py:2171-2178 (float32 host arithmetic; exact for any realistic map size)
  map_width = (max_x - min_x) + 2
  map_height = (max_z - min_z) + 2
  # shape (1,) like the reference's (the bbox is reduced over a (1, n) view, maps.py:2159-2169)
  width_offset = torch.tensor(map_width / 2., dtype=torch.float32) - torch.tensor([max_x + min_x]) / 2.
  height_offset = torch.tensor(map_height / 2., dtype=torch.float32) - torch.tensor([max_z + min_z]) / 2.
  tgt = nat.DmFuseTarget()
  tgt.Mh, tgt.Mw = map_height, map_width
  tgt.flip_h = bool(proj.flip_h)
  tgt.map_res = proj.map_res
  tgt.width_offset, tgt.height_offset = float(width_offset), float(height_offset)
  tgt.fill_value = get(fill_value, proj.fill_value, NINF)
  tgt.reduction = red
  topdown = _alloc_canvas((b, C, map_height, map_width), torch.float32, dev)
  mask = _alloc_canvas((b, C, map_height, map_width), torch.bool, dev)
  height = None if is_height_map else _alloc_canvas((b, C, map_height, map_width), torch.float32, dev)
  # the new map is in the global frame when the target is: pass 2 then leaves its box for the next merge
  track = bool(proj.to_global) and tgt.fill_value == tgt.fill_value and red < 2
  next_box = torch.empty((5,), dtype=torch.int64, device=dev) if track else None
  # ... and, whatever the frame, the per-plane rectangles of its valid cells (cell coordinates: no projector involved)
  track_planes = tgt.fill_value == tgt.fill_value and red < 2
  plane_box = torch.empty((b * C, 4), dtype=torch.int32, device=dev) if track_planes else None
  with torch.cuda.device(dev):
    rc = lib.dm_fuse_scatter_track_f32(sources, len(live), b, C, tgt, topdown.data_ptr(), mask.data_ptr(),
                                       nat.ptr(height), nat.ptr(next_box), nat.ptr(plane_box), nat.stream_ptr(dev))
  nat.check(rc, "dm_fuse_scatter_track_f32")
  new_proj = proj.clone(width_offset=width_offset, height_offset=height_offset, map_width=map_width,
                        map_height=map_height)
  out = TopdownMap(topdown_map=topdown, mask=mask, height_map=topdown if is_height_map else height,
                   map_projector=new_proj, is_height_map=is_height_map)
  if track or track_planes:
    out._tracked_box = _TrackedBox(next_box, mask, new_proj, plane_box)
  return out


class _TrackedBox():
  """Bounding box (device, 5 x int64) that pass 1 of fuse_topdown_maps would compute for a map this module wrote,
  with what it was derived from: the mask tensor (and its in-place version counter) and the projector's
  dequantisation parameters.  Anything that no longer matches makes the map an ordinary, scanned source again."""

  def __init__(self, box: Optional[torch.Tensor], mask: torch.Tensor, proj: MapProjector,
               plane_box: Optional[torch.Tensor] = None):
    self.box = box              # None: only the per-plane rectangles are tracked (local-frame target)
    self.plane_box = plane_box  # (b*C, 4) int32 device: rows [min, max], columns [min, max] of every plane's valid cells
    self.mask = mask
    self.mask_version = mask._version
    self.proj = proj
    self.key = self.proj_key(proj)

  @staticmethod
  def proj_key(proj: MapProjector):
    f = lambda x: None if x is None else tuple(float(v) for v in torch.as_tensor(x).reshape(-1))
    return (f(proj.width_offset), f(proj.height_offset), float(proj.map_res), bool(proj.flip_h), bool(proj.to_global))


def _tracked_box_valid(m: TopdownMap, target: MapProjector, dev: torch.device) -> bool:
  tb = getattr(m, "_tracked_box", None)
  if tb is None or tb.box is None or not target.to_global:
    return False
  return (m.mask is tb.mask and tb.mask._version == tb.mask_version and tb.box.device == dev
          and m.proj is tb.proj and tb.proj_key(m.proj) == tb.key and tb.key[4]
          and float(target.map_res) == tb.key[2])


def merge_into_canvas(world: TopdownMap, new_map: TopdownMap, canvas_shape: Tuple[int, int],
                      map_projector: MapProjector, fill_value: Optional[float] = None,
                      reduction: Optional[Reduction] = None) -> TopdownMap:
  """Opt-in fixed-canvas merge (not in the reference; its closest relative is project(..., canvas=,
  canvas_masks=), maps.py:1089-1173): the valid cells of `new_map` become points again exactly as in
  fuse_topdown_maps (dequantise, local → global with the map's own pose), are quantised with the
  canvases' FIXED offsets (map_width / 2, map_height / 2: world origin at the centre) and max-merged in
  place into `world`'s tensors, which are allocated on the first call.  One launch, no host sync.
  `world`'s tensors are updated in place and shared with the returned map."""
  Hc, Wc = int(canvas_shape[0]), int(canvas_shape[1])
  red = utils._reduction_code(get(reduction, map_projector.reduction), fused=False)
  if new_map.is_empty:
    return world
  dev = _pick_device(map_projector.device, new_map.mask, new_map.height_map)
  b, C, _, _ = utils.to_4D_image(new_map.mask).shape
  is_height_map = bool(new_map.is_height_map)
  fill = get(fill_value, map_projector.fill_value, NINF)
  lib = nat.lib()
  if world is None or world.is_empty:
    topdown = torch.empty((b, C, Hc, Wc), dtype=torch.float32, device=dev)
    mask = torch.empty((b, C, Hc, Wc), dtype=torch.bool, device=dev)
    height = None if is_height_map else torch.empty_like(topdown)
    with torch.cuda.device(dev):
      rc = lib.dm_fuse_canvas_init_f32(topdown.data_ptr(), mask.data_ptr(), nat.ptr(height), topdown.numel(),
                                       fill, nat.stream_ptr(dev))
    nat.check(rc, "dm_fuse_canvas_init_f32")
  else:
    assert bool(world.is_height_map) == is_height_map, "All maps must be the same type of maps"
    topdown, mask = world.topdown_map, world.mask
    height = None if is_height_map else world.height_map
    assert tuple(topdown.shape) == (b, C, Hc, Wc), f"world canvas {tuple(topdown.shape)} != {(b, C, Hc, Wc)}"
  target = map_projector.clone(to_global=True, width_offset=Wc / 2., height_offset=Hc / 2., map_width=Wc,
                               map_height=Hc)
  keep: list = []
  h, w = new_map.mask.shape[-2:]
  sources = (nat.DmFuseSource * 1)(_fuse_source(new_map, target, b, C, C * h * w, dev, keep))
  tgt = nat.DmFuseTarget()
  tgt.Mh, tgt.Mw = Hc, Wc
  tgt.flip_h = bool(target.flip_h)
  tgt.map_res = target.map_res
  tgt.width_offset, tgt.height_offset = Wc / 2., Hc / 2.
  tgt.fill_value = fill
  tgt.reduction = red
  with torch.cuda.device(dev):
    rc = lib.dm_fuse_inplace_f32(sources, 1, b, C, tgt, topdown.data_ptr(), mask.data_ptr(), nat.ptr(height),
                                 nat.stream_ptr(dev))
  nat.check(rc, "dm_fuse_inplace_f32")
  return TopdownMap(topdown_map=topdown, mask=mask, height_map=topdown if is_height_map else height,
                    map_projector=target, is_height_map=is_height_map)


# ======== MapBuilder (maps.py:2289-2550) ==========================================================

class _NativeBuilder():
  """Owner of one DmBuilder handle (csrc/dm_builder.cu): a MapBuilder's step configuration frozen into C."""

  def __init__(self, proj: MapProjector, key: tuple, dev: torch.device):
    (b, H, W, _, plot_global, woff, hoff, Mw, Mh, pitch, camh, res, _, _, tdmin, tdmax, thmax, clip, flip, fill,
     red) = key
    cfg = nat.DmBuilderCfg()
    p = cfg.proj
    p.H, p.W, p.C, p.Mh, p.Mw = H, W, 0, Mh, Mw
    k = proj.cam_params
    p.fx, p.fy, p.cx, p.cy = k.fx, k.fy, k.cx, k.cy
    p.map_res = res
    p.has_trunc_depth_min, p.has_trunc_depth_max = tdmin is not None, tdmax is not None
    p.has_trunc_height_max = thmax is not None
    p.trunc_depth_min, p.trunc_depth_max, p.trunc_height_max = tdmin or 0., tdmax or 0., thmax or 0.
    p.clip_border = int(clip) if clip is not None else 0
    p.flip_h = flip
    p.fill_value = 0. if fill is None else fill      # utils.py:472-473
    p.reduction = red
    cfg.b, cfg.plot_to_global = b, plot_global
    R = prm.rotation_matrices([1., 0., 0.], torch.tensor([pitch], dtype=torch.float32))  # maps.py:789-793
    cfg.pitch_R[:] = [float(v) for v in R.reshape(-1)]
    cfg.cam_height = camh
    cfg.width_offset, cfg.height_offset = woff, hoff
    skew, skew_sq, _ = prm._skew_terms([0., 1., 0.])
    cfg.yaw_skew[:] = [float(v) for v in skew.reshape(-1)]
    cfg.yaw_skew_sq[:] = [float(v) for v in skew_sq.reshape(-1)]
    self.fill = get(fill, NINF)                       # maps.py:2246: get(fill_value, proj.fill_value, NINF)
    cfg.merge_fill_value = self.fill
    cfg.merge_reduction = red
    self._lib = nat.lib()
    handle = nat.ctypes.c_void_p()
    with torch.cuda.device(dev):
      nat.check(self._lib.dm_builder_create(cfg, dev.index, nat.ctypes.byref(handle)), "dm_builder_create")
    self.handle = handle

  def __del__(self):
    try:
      if self.handle:
        self._lib.dm_builder_destroy(self.handle)
        self.handle = None
    except Exception:
      pass


class MapBuilder():
  """Plots a local top-down map per frame and merges it into a growing world map."""

  def __init__(self, map_projector: MapProjector, world_map: Optional[TopdownMap] = None,
               fixed_canvas: Optional[Tuple[int, int]] = None, native_step: bool = True):
    """`fixed_canvas=(map_height, map_width)` opts into the in-place world map (not in the reference,
    SURVEY.md §8f-2): global canvases of that size are allocated at the first merge, the world origin sits
    at their centre, and every merge max-merges the new map's cells into them with one kernel launch —
    no bounding box, no host sync, no reallocation.  Points that fall outside the canvas are dropped.
    `native_step`: height-map steps of a global-frame builder (value_map / valid_map None, CenterMode.none, device
    depth tensors) run with their host side in C (dm_builder_*, csrc/dm_builder.cu) — same kernels, same results,
    two library calls per step instead of ~1000 interpreter calls; anything else takes the general path."""
    self._proj = map_projector
    self._world_map = world_map if world_map is not None else TopdownMap(map_projector=self.proj.clone())
    self._fixed = None if fixed_canvas is None else (int(fixed_canvas[0]), int(fixed_canvas[1]))
    self._native = bool(native_step)
    self._handles: Dict[tuple, "_NativeBuilder"] = {}

  @property
  def proj(self) -> MapProjector:
    return self._proj

  @property
  def world_map(self) -> TopdownMap:
    return self._world_map

  def reset(self, depth_map: Optional[np.ndarray] = None, value_map: Optional[np.ndarray] = None,
            valid_map: Optional[np.ndarray] = None, cam_pose: Optional[np.ndarray] = None,
            center_mode: CenterMode = CenterMode.none, **kwargs):
    """Forget the world map; if a frame is given, plot and merge it (maps.py:2312-2355)."""
    self._world_map = TopdownMap(map_projector=self.proj.clone())
    if depth_map is None:
      return None
    return self.step(depth_map=depth_map, value_map=value_map, valid_map=valid_map, cam_pose=cam_pose,
                     center_mode=center_mode, **kwargs)

  def step(self, depth_map: np.ndarray, value_map: Optional[np.ndarray] = None,
           valid_map: Optional[np.ndarray] = None, cam_pose: Optional[np.ndarray] = None,
           center_mode: CenterMode = CenterMode.none, merge: bool = True, keep_pose: bool = False,
           **kwargs: Dict[str, Any]) -> TopdownMap:
    """Plot the frame's local map and (by default) merge it into the world map
    (maps.py:2357-2406).  Returns the local map."""
    if self._native and merge and not keep_pose and value_map is None and valid_map is None:
      done = self._native_step(depth_map, cam_pose, center_mode, kwargs)
      if done is not None:
        return done
    topdown_map = self.plot(depth_map=depth_map, value_map=value_map, valid_map=valid_map, cam_pose=cam_pose,
                            center_mode=center_mode, **kwargs)
    if merge:
      self.merge(topdown_map, keep_pose=keep_pose)
    return topdown_map

  def plot(self, depth_map: np.ndarray, value_map: Optional[np.ndarray] = None,
           valid_map: Optional[np.ndarray] = None, cam_pose: Optional[np.ndarray] = None,
           center_mode: CenterMode = CenterMode.none, **kwargs: Dict[str, Any]) -> TopdownMap:
    """orth_project with get_height_map=True, wrapped as a TopdownMap that remembers the pose and
    offsets it was plotted with (maps.py:2408-2469)."""
    cam_pose = get(cam_pose, self.proj.cam_pose, np.array([0., 0., 0.], dtype=np.float32))
    width_offset, height_offset = self._compute_offsets(cam_pose=cam_pose, center_mode=center_mode, **kwargs)
    kwargs['width_offset'] = width_offset
    kwargs['height_offset'] = height_offset
    kwargs.pop('get_height_map', None)
    topdown_map, mask, height_map = self.proj.orth_project(
      depth_map=depth_map, value_map=value_map, valid_map=valid_map, cam_pose=cam_pose,
      get_height_map=True, **kwargs)
    clone_kwargs = {k: v for k, v in kwargs.items() if k in _CTOR_ARGS}
    return TopdownMap(topdown_map=topdown_map, mask=mask, height_map=height_map,
                      map_projector=self.proj.clone(cam_pose=cam_pose, **clone_kwargs),
                      is_height_map=(value_map is None and kwargs.get('label_map') is None))

  def merge(self, topdown_map: TopdownMap, keep_pose: bool = False, fill_value: Optional[float] = None,
            reduction: Optional[Reduction] = None) -> TopdownMap:
    """Fuse `topdown_map` into the world map, in the new map's camera frame unless `keep_pose`
    (maps.py:2471-2508)."""
    if self._world_map is None:
      self._world_map = TopdownMap(map_projector=self.proj.clone())
    cam_pose = self._world_map.proj.cam_pose if keep_pose else topdown_map.proj.cam_pose
    if self._fixed is not None:
      self._world_map = merge_into_canvas(self._world_map, topdown_map, canvas_shape=self._fixed,
                                          map_projector=self.proj.clone(cam_pose=cam_pose),
                                          fill_value=fill_value, reduction=reduction)
      return self._world_map
    self._world_map = fuse_topdown_maps(self._world_map, topdown_map,
                                        map_projector=self.proj.clone(cam_pose=cam_pose),
                                        fill_value=fill_value, reduction=reduction)
    return self._world_map

  # ---- the step with its host side in C (csrc/dm_builder.cu) ----------------------------------------------------

  _NATIVE_KW = frozenset(("to_global", "width_offset", "height_offset", "map_width", "map_height"))

  def _native_step(self, depth_map, cam_pose, center_mode, kwargs) -> Optional[TopdownMap]:
    """MapBuilder.step for the case dm_builder_* covers; None when the call needs the general path."""
    proj = self.proj
    if (CenterMode(center_mode) is not CenterMode.none or not set(kwargs) <= self._NATIVE_KW or not proj.to_global
        or not torch.is_tensor(depth_map) or not depth_map.is_cuda or depth_map.dtype is not torch.float32
        or depth_map.dim() != 4 or depth_map.shape[1] != 1 or not depth_map.is_contiguous()):
      return None
    get_kw = lambda name: get(kwargs.get(name), getattr(proj, name))
    scalars = [proj.cam_pitch, proj.cam_height, proj.map_res, get_kw("width_offset"), get_kw("height_offset"),
               get_kw("map_width"), get_kw("map_height")]
    if any(v is None or torch.is_tensor(v) or isinstance(v, (np.ndarray, list, tuple)) for v in scalars):
      return None
    try:
      red = utils._reduction_code(proj.reduction)
    except NotImplementedError:
      return None
    dev = depth_map.device
    b, _, H, W = depth_map.shape
    if (H, W) != (int(proj.height), int(proj.width)):
      return None
    world = self._world_map
    have_world = world is not None and not world.is_empty
    if have_world and not self._native_world_ok(world, b, dev):
      return None
    cam_pose = get(cam_pose, proj.cam_pose, np.array([0., 0., 0.], dtype=np.float32))
    plot_global = bool(get_kw("to_global"))
    woff, hoff = float(scalars[3]), float(scalars[4])
    Mw, Mh = int(scalars[5]), int(scalars[6])
    fill = proj.fill_value
    key = (b, H, W, dev.index, plot_global, woff, hoff, Mw, Mh, float(proj.cam_pitch), float(proj.cam_height),
           float(proj.map_res), proj.hfov, proj.vfov, proj.trunc_depth_min, proj.trunc_depth_max,
           proj.trunc_height_max, proj.clip_border, bool(proj.flip_h), fill, red)
    nb = self._handles.get(key)
    if nb is None:
      nb = self._handles[key] = _NativeBuilder(proj, key, dev)
    pose = prm.per_sample(cam_pose, b, (3,), "cam_pose").contiguous()
    # utils.py:325-326 with the reference's own torch-CPU ops (the last ulp of sin / cos matters)
    sin, cos = prm.yaw_sin_cos(pose)
    local_top = torch.empty((b, 1, Mh, Mw), dtype=torch.float32, device=dev)
    local_mask = torch.empty((b, 1, Mh, Mw), dtype=torch.bool, device=dev)
    lib = nat.lib()
    stream = nat.stream_ptr(dev)
    def make_local():
      """The local map as plot() describes it (maps.py:2459-2469); offsets as compute_center_offsets returns them.
      Built AFTER the step's kernels are queued: host bookkeeping belongs behind GPU work, not in front of it."""
      local_kw = {k: v for k, v in kwargs.items() if k in _CTOR_ARGS}
      local_kw["width_offset"] = torch.tensor([woff], dtype=torch.float32)
      local_kw["height_offset"] = torch.tensor([hoff], dtype=torch.float32)
      return TopdownMap(topdown_map=local_top, mask=local_mask, height_map=local_top,
                        map_projector=proj.clone(cam_pose=cam_pose, **local_kw), is_height_map=True)
    target = proj.clone(cam_pose=cam_pose)
    with torch.cuda.device(dev):
      if self._fixed is not None:
        Hc, Wc = self._fixed
        if not have_world:
          topdown = torch.empty((b, 1, Hc, Wc), dtype=torch.float32, device=dev)
          mask = torch.empty((b, 1, Hc, Wc), dtype=torch.bool, device=dev)
          nat.check(lib.dm_fuse_canvas_init_f32(topdown.data_ptr(), mask.data_ptr(), None, topdown.numel(),
                                                get(fill, NINF), stream), "dm_fuse_canvas_init_f32")
        else:
          topdown, mask = world.topdown_map, world.mask
        canvas = nat.DmMapRef(topdown.data_ptr(), mask.data_ptr(), Hc, Wc, Wc / 2., Hc / 2., None, None)
        nat.check(lib.dm_builder_step_fixed(nb.handle, depth_map.data_ptr(), pose.data_ptr(), sin.data_ptr(),
                                            cos.data_ptr(), local_top.data_ptr(), local_mask.data_ptr(), canvas,
                                            stream), "dm_builder_step_fixed")
        self._world_map = TopdownMap(
          topdown_map=topdown, mask=mask, height_map=topdown, is_height_map=True,
          map_projector=target.clone(to_global=True, width_offset=Wc / 2., height_offset=Hc / 2., map_width=Wc,
                                     map_height=Hc))
        return make_local()
      wref = None
      if have_world:
        box = world._tracked_box.box if _tracked_box_valid(world, target, dev) else None
        tb = getattr(world, "_tracked_box", None)
        planes = None
        if (tb is not None and tb.plane_box is not None and world.mask is tb.mask
            and tb.mask._version == tb.mask_version and tb.plane_box.device == dev):
          planes = tb.plane_box
        wref = nat.DmMapRef(world.topdown_map.data_ptr(), world.mask.data_ptr(), world.mask.shape[-2],
                            world.mask.shape[-1], float(world.proj.width_offset), float(world.proj.height_offset),
                            nat.ptr(box), nat.ptr(planes))
      shape = nat.DmMergeShape()
      # Everything the merge needs that does not depend on the bounding box is set up BEFORE the call that waits for it:
      # the GPU idles from the moment the box arrives until the merge kernels are queued.  The new canvas is as a rule
      # in the size class of the old one (the map grows by a few cells per step), so canvases of that class are taken
      # from the allocator ahead of the sync and only replaced when the box asks for more.
      track = nb.fill == nb.fill
      next_box = torch.empty((5,), dtype=torch.int64, device=dev) if track else None
      next_planes = torch.empty((b, 4), dtype=torch.int32, device=dev) if track else None
      spec_top = spec_mask = None
      if have_world and world.mask.numel() >= (1 << 18):
        cap = _canvas_cap(world.mask.numel())
        spec_top = torch.empty((cap,), dtype=torch.float32, device=dev)
        spec_mask = torch.empty((cap,), dtype=torch.bool, device=dev)
      # ... and their fill is queued behind the box's copy: it runs while the host waits for the box
      # (prefilled: the old map's cells + 2 % — the map grows by a row or a column now and then; dm_builder_merge fills
      # whatever tail the new map has beyond that, and the rest of the size class stays untouched)
      n_guess = 0 if spec_top is None else min(spec_top.numel(), world.mask.numel() + world.mask.numel() // 50 + 4096)
      nat.check(lib.dm_builder_plot_prefill(nb.handle, depth_map.data_ptr(), pose.data_ptr(), sin.data_ptr(),
                                            cos.data_ptr(), local_top.data_ptr(), local_mask.data_ptr(), wref, None,
                                            nat.ptr(spec_top), nat.ptr(spec_mask), n_guess, stream),
                "dm_builder_plot_prefill")
      local = make_local()  # while the projection, the box reduction and the prefill run
      nat.check(lib.dm_builder_plot_wait(nb.handle, shape), "dm_builder_plot_wait")
      if shape.n_valid == 0:  # maps.py:2217-2225
        self._world_map = TopdownMap(topdown_map=local.topdown_map, mask=local.mask, height_map=local.height_map,
                                     map_projector=target)
        return local
      mh, mw = shape.map_height, shape.map_width
      n_new = b * mh * mw
      if spec_top is not None and (1 << 18) <= n_new <= spec_top.numel():
        topdown = spec_top[:n_new].view(b, 1, mh, mw)
        mask = spec_mask[:n_new].view(b, 1, mh, mw)
      else:
        topdown = _alloc_canvas((b, 1, mh, mw), torch.float32, dev)
        mask = _alloc_canvas((b, 1, mh, mw), torch.bool, dev)
      spec_top = spec_mask = None
      out = nat.DmMapRef(topdown.data_ptr(), mask.data_ptr(), mh, mw, shape.width_offset, shape.height_offset,
                         nat.ptr(next_box), nat.ptr(next_planes))
      nat.check(lib.dm_builder_merge(nb.handle, out, stream), "dm_builder_merge")
    new_proj = target.clone(width_offset=torch.tensor([shape.width_offset], dtype=torch.float32),
                            height_offset=torch.tensor([shape.height_offset], dtype=torch.float32),
                            map_width=mw, map_height=mh)
    merged = TopdownMap(topdown_map=topdown, mask=mask, height_map=topdown, map_projector=new_proj,
                        is_height_map=True)
    if track:
      merged._tracked_box = _TrackedBox(next_box, mask, new_proj, next_planes)
    # (the scatter pass still reads the old world map and the local map on the stream: torch's caching allocator
    # hands freed blocks to later work of the same stream only, so dropping the old map here is safe)
    self._world_map = merged
    return local

  def _native_world_ok(self, world: TopdownMap, b: int, dev: torch.device) -> bool:
    """The world map is one the C step can take as it is: a contiguous (b, 1, h, w) float32 height map + bool mask on
    `dev`, in the global frame, with this builder's resolution / flip and scalar offsets."""
    p, proj = world.proj, self.proj
    top, mask = world.topdown_map, world.mask
    if not (world.is_height_map and p is not None and p.to_global and torch.is_tensor(top) and torch.is_tensor(mask)):
      return False
    if self._fixed is not None and tuple(top.shape[-2:]) != self._fixed:
      return False
    ok_t = lambda t, dt: (t.device == dev and t.dtype is dt and t.dim() == 4 and t.shape[0] == b and t.shape[1] == 1
                          and t.is_contiguous())
    if not (ok_t(top, torch.float32) and ok_t(mask, torch.bool) and top.shape == mask.shape):
      return False
    def one(v):
      if v is None:
        return False
      if torch.is_tensor(v):
        return v.numel() == 1
      return np.asarray(v).size == 1
    return (float(p.map_res) == float(proj.map_res) and bool(p.flip_h) == bool(proj.flip_h)
            and one(p.width_offset) and one(p.height_offset))

  def _compute_offsets(self, cam_pose: np.ndarray, width_offset: Optional[np.ndarray] = None,
                       height_offset: Optional[np.ndarray] = None, map_res: Optional[float] = None,
                       map_width: Optional[int] = None, map_height: Optional[int] = None,
                       to_global: Optional[bool] = None, center_mode: Optional[CenterMode] = None,
                       **kw_) -> Tuple[torch.Tensor, torch.Tensor]:
    return self.proj.compute_center_offsets(cam_pose=cam_pose, width_offset=width_offset,
                                            height_offset=height_offset, map_res=map_res, map_width=map_width,
                                            map_height=map_height, to_global=to_global, center_mode=center_mode)
