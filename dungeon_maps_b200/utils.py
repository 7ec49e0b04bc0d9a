"""Drop-in for `dungeon_maps.utils` (reference: /root/reference/dungeon_maps/utils.py).

Same public names, argument meaning and error behaviour; the arithmetic runs
in the sm_100a kernels behind include/dungeon_maps_b200.h.  Only tensor/shape
plumbing (conversion, reshape, index raveling on integer tensors) is torch.
Differences from the reference, all consequences of "CUDA only, no fallback":
  * tensors that are not on a CUDA device are moved to one (`device=` or the
    current CUDA device) and results live there;
  * batch > 1 works (the reference raises, utils.py:311-316);
  * Reduction.sum / mean / prod fold the hits of a cell in the reference's CPU order (ascending point index,
    csrc/dm_ordered.cu): bit-identical to the reference's CPU path and deterministic; max / min are atomics.
"""
import enum
from dataclasses import dataclass
from typing import Any, Optional, Tuple, Union

import numpy as np
import torch

from . import _native as nat
from . import _params as prm

__all__ = [
  'NINF', 'Reduction', 'CameraIntrinsics', 'get_camera_intrinsics',
  'to_numpy', 'to_tensor', 'to_tensor_like', 'validate_tensors',
  'translate', 'rotate', 'ravel_index', 'scatter_tensor',
  'to_4D_image', 'from_4D_image', 'generate_image_coords', 'generate_crop_grid', 'image_sample',
]

NINF = -np.inf
ANGLE_EPS = prm.ANGLE_EPS
Float3D = Tuple[float, float, float]


@enum.unique
class Reduction(str, enum.Enum):
  """utils.py:52-67; Reduction(None) is Reduction.max."""
  max = 'max'
  min = 'min'
  sum = 'sum'
  mean = 'mean'
  prod = 'prod'

  @classmethod
  def _missing_(cls, value):
    if value is None:
      return cls.max


_RED_CODES = {Reduction.max: 0, Reduction.min: 1, Reduction.sum: 2, Reduction.mean: 3, Reduction.prod: 4}


def _reduction_code(reduction, fused: bool = True) -> int:
  """Kernel code of a Reduction.  The single-pass projection kernels implement max and min (fused=True raises for the
  others: orth_project then takes its composed path); scatter_tensor / project / fuse_topdown_maps /
  merge_into_canvas implement all five (fused=False)."""
  red = Reduction(reduction)
  code = _RED_CODES[red]
  if fused and code > 1:
    raise NotImplementedError(
      f"Reduction.{red.value} is not implemented by the single-pass projection kernel (max and min are); "
      "orth_project / scatter_tensor / project / fuse_topdown_maps handle it through the ordered scatter")
  return code


@dataclass
class CameraIntrinsics:
  """utils.py:79-92."""
  cx: float
  cy: float
  fx: float
  fy: float


def get_camera_intrinsics(width: float, height: float, hfov: float,
                          vfov: Optional[float] = None) -> CameraIntrinsics:
  """Pinhole intrinsics from the field of view, in float64 like utils.py:94-116."""
  cx = width / 2.
  cy = height / 2.
  fx = cx / np.tan(hfov / 2.)
  fy = cy / np.tan(vfov / 2.) if vfov is not None else fx
  return CameraIntrinsics(cx=cx, cy=cy, fx=fx, fy=fy)


# ---- tensor plumbing (utils.py:119-227) ---------------------------------------------------

def to_numpy(inputs: Any, dtype: Optional[np.dtype] = None) -> np.ndarray:
  arr = inputs.detach().cpu().numpy() if torch.is_tensor(inputs) else np.asarray(inputs)
  return arr.astype(dtype=dtype or arr.dtype)


def to_tensor(inputs: Any, dtype: Optional[torch.dtype] = None,
              device: Optional[torch.device] = None, **kwargs) -> torch.Tensor:
  if torch.is_tensor(inputs):
    t = inputs
  elif isinstance(inputs, np.ndarray):
    t = torch.from_numpy(inputs)
  else:
    t = torch.tensor(inputs, dtype=dtype)
  return t.to(device=device, dtype=dtype, **kwargs)


def to_tensor_like(inputs: Any, tensor: torch.Tensor) -> torch.Tensor:
  assert torch.is_tensor(tensor), f"`tensor` must be a torch.Tensor, got {type(tensor)}"
  return to_tensor(inputs, dtype=tensor.dtype, device=tensor.device)


def validate_tensors(*args: Any, same_device: Optional[Union[bool, torch.device]] = None,
                     same_dtype: Optional[Union[bool, torch.dtype]] = None, keep_tuple: bool = False):
  """Converts every argument to a tensor; `same_device=True` / `same_dtype=True` follow the
  first argument (utils.py:182-227)."""
  if not args:
    return
  first = to_tensor(args[0])
  if isinstance(same_device, bool):
    same_device = first.device if same_device else None
  if isinstance(same_dtype, bool):
    same_dtype = first.dtype if same_dtype else None
  out = tuple(to_tensor(a, device=same_device, dtype=same_dtype) for a in args)
  return out[0] if (len(out) == 1 and not keep_tuple) else out


def _on_cuda(x: Any, dtype: torch.dtype, device=None) -> torch.Tensor:
  """Contiguous tensor of `dtype` on the CUDA device the kernels will use."""
  if torch.is_tensor(x) and x.is_cuda and device is None:
    dev = x.device
  else:
    dev = nat.require_cuda(device)
  t = to_tensor(x)
  return t.to(device=dev, dtype=dtype).contiguous()


def _device_of(*tensors, device=None) -> torch.device:
  if device is not None:
    return nat.require_cuda(device)
  for t in tensors:
    if torch.is_tensor(t) and t.is_cuda:
      return t.device
  return nat.require_cuda(None)


# ---- transformations ----------------------------------------------------------------------

def _apply_steps(points: torch.Tensor, steps_host: torch.Tensor, n_steps: int) -> torch.Tensor:
  """points (b, n, 3) cuda f32; steps_host (b, n_steps*16) → dm_transform_points_f32."""
  b, n = points.shape[0], points.shape[1]
  dev = points.device
  steps = prm.upload(steps_host, dev)
  out = torch.empty_like(points)
  with torch.cuda.device(dev):
    rc = nat.lib().dm_transform_points_f32(points.data_ptr(), steps.data_ptr(), n_steps, b, n,
                                           out.data_ptr(), nat.stream_ptr(dev))
  nat.check(rc, "dm_transform_points_f32")
  return out


def translate(points: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
  """points (b, ..., 3) + offsets (b, 3)  (utils.py:229-259)."""
  pts = _on_cuda(points, torch.float32)
  shape = pts.shape
  off = prm.host_f32(offsets, (3,))
  b = shape[0] if pts.dim() > 1 else 1
  flat = pts.reshape(b, -1, 3)
  off = prm.per_sample(off, b, (3,), "offsets")
  steps = prm.pack_steps(prm.STEP_ADD, None, off, b, 1 << 30)
  return _apply_steps(flat, steps, 1).reshape(shape)


def rotate(points: torch.Tensor, axis: torch.Tensor, angle: torch.Tensor,
           angle_eps: float = ANGLE_EPS) -> torch.Tensor:
  """Rodrigues rotation of points (b, ..., 3) about `axis` (b, 3) by `angle` (b,)
  (utils.py:261-330).  The matrix is built on the host, the product runs on the GPU."""
  pts = _on_cuda(points, torch.float32)
  shape = pts.shape
  b = shape[0]
  flat = pts.reshape(b, -1, 3)
  ang = prm.per_sample(angle, b, (), "angle")
  R = prm.rotation_matrices(prm.host_f32(axis, (3,)), ang, angle_eps)
  steps = prm.pack_steps(prm.STEP_ROT, R, None, b, flat.shape[1])
  return _apply_steps(flat, steps, 1).reshape(shape)


def ravel_index(index: torch.Tensor, shape: torch.Size, keepdim: bool = False) -> torch.Tensor:
  """Row-major flattening of (..., n) indices, like np.ravel_multi_index (utils.py:332-370).
  Integer plumbing on whatever device `index` lives on."""
  index = to_tensor(index, dtype=torch.int64)
  dims = torch.tensor((1,) + tuple(shape)[::-1], dtype=torch.int64, device=index.device)
  strides = torch.cumprod(dims, dim=0)[:-1].flip(0)
  return (index * strides).sum(dim=-1, keepdim=keepdim)


def scatter_tensor(canvas: torch.Tensor, indices: torch.Tensor, values: torch.Tensor,
                   masks: Optional[torch.Tensor] = None, fill_value: Optional[float] = None,
                   reduction: Optional[Union[str, Reduction]] = None,
                   _validate_args: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
  """Scatter-reduce `values` (b..., N) into an n-D `canvas` (b..., d1..dn) at `indices`
  (b..., N, n); returns the new canvas and the "cell changed" mask (utils.py:389-492)."""
  red = _reduction_code(reduction, fused=False)
  dev = _device_of(canvas, indices, values)
  canvas = to_tensor(canvas).to(device=dev, dtype=torch.float32)
  indices = to_tensor(indices).to(device=dev, dtype=torch.int64)
  values = to_tensor(values).to(device=dev, dtype=torch.float32)
  n = indices.shape[-1]
  assert canvas.dim() > n, f"The rank of `canvas` must be greater than {n}, got {canvas.dim()}"
  dims = tuple(canvas.shape[-n:])
  batch_dims = tuple(canvas.shape[:-n])
  N = values.shape[-1]
  full = torch.broadcast_shapes(batch_dims + (N,), tuple(values.shape), tuple(indices.shape[:-1]))
  values = values.expand(full).contiguous()
  indices = indices.expand(full + (n,))
  valid = None
  if masks is not None:
    valid = to_tensor(masks).to(device=dev, dtype=torch.bool).expand(full)
  # bounds per dimension, then ravel everything but the last dim into the row (utils.py:448-460)
  inb = ((indices >= 0) & (indices < torch.tensor(dims, device=dev))).all(dim=-1)
  valid = inb if valid is None else (valid & inb)
  if n == 2:
    rows, cols, Mh, Mw = indices[..., 0], indices[..., 1], dims[0], dims[1]
  else:
    flat = ravel_index(torch.where(valid.unsqueeze(-1), indices, torch.zeros_like(indices)), dims)
    rows, cols, Mh, Mw = torch.zeros_like(flat), flat, 1, int(np.prod(dims))
  coords = torch.stack((rows, cols), dim=-1).contiguous()
  valid_u8 = valid.contiguous()
  B = int(np.prod(full[:-1])) if len(full) > 1 else 1
  canvas_in = canvas.expand(full[:-1] + dims).contiguous()
  out = torch.empty_like(canvas_in)
  mask = torch.empty(canvas_in.shape, dtype=torch.bool, device=dev)
  with torch.cuda.device(dev):
    rc = nat.lib().dm_scatter_f32(values.data_ptr(), coords.data_ptr(), valid_u8.data_ptr(), B, N, Mh, Mw,
                                  int(fill_value is not None), float(0. if fill_value is None else fill_value),
                                  red, canvas_in.data_ptr(), out.data_ptr(), mask.data_ptr(),
                                  nat.stream_ptr(dev))
  nat.check(rc, "dm_scatter_f32")
  return out, mask


# ---- image helpers (utils.py:494-652) -------------------------------------------------------

def to_4D_image(image: torch.Tensor) -> torch.Tensor:
  nd = image.dim()
  assert nd in [2, 3, 4], f"only supports 2/3/4D images while {nd}-D are given."
  return image[(None,) * (4 - nd)]


def from_4D_image(image: torch.Tensor, ndims: int) -> torch.Tensor:
  assert image.dim() == 4, f"`image` must be a 4D tensor, while {image.dim()}-D are given."
  if ndims == 2:
    return image[0, 0]
  if ndims == 3:
    return image[0]
  return image


def generate_image_coords(image_shape: torch.Size, dtype: Optional[torch.dtype] = None,
                          device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
  """Broadcast views of the column / row index of every pixel (utils.py:535-569)."""
  dtype = dtype or torch.float32
  nd = len(image_shape)
  if nd < 2:
    raise ValueError(f"rank of `image_shape` must be at east 2D, got {nd}")
  h, w = image_shape[-2], image_shape[-1]
  lead = (1,) * (nd - 2)
  x = torch.arange(w, dtype=dtype, device=device).view(lead + (1, w)).expand(tuple(image_shape))
  y = torch.arange(h, dtype=dtype, device=device).view(lead + (h, 1)).expand(tuple(image_shape))
  return x, y


def generate_crop_grid(center: torch.Tensor, image_width: int, image_height: int, crop_width: int,
                       crop_height: int, device: Optional[torch.device] = None) -> torch.Tensor:
  """Normalised sampling grid of a crop around `center` on the 1-px padded image
  (utils.py:571-611).  Divisions use tensor divisors so CUDA rounds like the CPU reference."""
  center = to_tensor(center, device=device).view(-1, 2).to(dtype=torch.float32)
  dev = center.device
  b = center.shape[0]
  center = center + 1
  h, w = image_height + 2, image_width + 2
  x, y = generate_image_coords((b, crop_height, crop_width), dtype=torch.float32, device=dev)
  cx = (center[:, 0] - w / 2.).view(-1, 1, 1)
  cy = (center[:, 1] - h / 2.).view(-1, 1, 1)
  half_w = torch.tensor(w / 2., dtype=torch.float32, device=dev)
  half_h = torch.tensor(h / 2., dtype=torch.float32, device=dev)
  gx = torch.div(x - crop_width / 2. + cx, half_w)
  gy = torch.div(y - crop_height / 2. + cy, half_h)
  return torch.stack((gx, gy), dim=-1)


def image_sample(image: torch.Tensor, grid: torch.Tensor, fill_value: Optional[float] = None,
                 mode: str = 'nearest', _validate_args: bool = True) -> torch.Tensor:
  """Generic grid sampling of a padded image (utils.py:613-652).  Off the hot path: the crop
  used by TopdownMap.select is the fused dm_crop_nearest kernel; this general entry keeps
  the reference's behaviour for arbitrary grids through torch's grid_sample."""
  image, grid = validate_tensors(image, grid, same_device=True)
  image = to_4D_image(image)
  padding_mode = 'border'
  if fill_value is None:
    fill_value, padding_mode = 0.0, 'zeros'
  orig = image.dtype
  image = torch.nn.functional.pad(image.to(grid.dtype), [1, 1, 1, 1], mode='constant', value=float(fill_value))
  out = torch.nn.functional.grid_sample(image, grid, mode=mode, padding_mode=padding_mode, align_corners=True)
  return out.to(dtype=orig)
