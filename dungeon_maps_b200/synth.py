"""Synthetic depth / semantic frames (stand-in for the reference's `sim` package,
dungeon_maps/sim/dungeon.py:178-210, which needs moderngl + EGL).

Everything is a pure function of (shape, seed): a counter-based integer hash
(splitmix64 on the element index) instead of a library RNG, so the same
frames come out on any device, torch version or machine — tests, golden
fixtures and bench.py all draw from here.
"""
import math
from typing import Optional, Sequence, Tuple

import torch

_M64 = (1 << 64) - 1


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
  """Logical shift right on int64 (torch's >> is arithmetic)."""
  return (x >> k) & ((1 << (64 - k)) - 1)


def _i64(v: int) -> int:
  v &= _M64
  return v - (1 << 64) if v >= (1 << 63) else v


def hash_u24(n: int, seed: int, device=None, offset: int = 0) -> torch.Tensor:
  """n pseudo-random integers in [0, 2^24), splitmix64 of (index + offset, seed)."""
  i = torch.arange(offset, offset + n, dtype=torch.int64, device=device)
  z = i * _i64(0x9E3779B97F4A7C15) + _i64(seed * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D)
  z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
  z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
  z = z ^ _lsr(z, 31)
  return _lsr(z, 40)


def uniform(shape: Sequence[int], seed: int, lo: float = 0.0, hi: float = 1.0,
            device=None) -> torch.Tensor:
  """float32 tensor ~ U[lo, hi) (24-bit mantissa grid, exact on every backend)."""
  n = int(math.prod(shape))
  u = hash_u24(n, seed, device).to(torch.float32) * (1.0 / (1 << 24))
  return (u * (hi - lo) + lo).reshape(tuple(shape))


def iid_depth(b: int, H: int, W: int, seed: int = 0, lo: float = 0.1, hi: float = 10.0,
              device=None) -> torch.Tensor:
  """(b,1,H,W) i.i.d. U[lo,hi) depth: worst-case locality."""
  return uniform((b, 1, H, W), seed, lo, hi, device)


def poses(b: int, seed: int = 0, xz: float = 1.0, yaw: float = math.pi, device=None) -> torch.Tensor:
  """(b,3) [x, z, yaw] with x,z ~ U(-xz,xz), yaw ~ U(-yaw,yaw)."""
  u = uniform((b, 3), seed ^ 0x5EED, -1.0, 1.0, device)
  return u * torch.tensor([xz, xz, yaw], dtype=torch.float32, device=device)


def block_labels(b: int, C: int, H: int, W: int, seed: int = 0, block: int = 16, device=None) -> torch.Tensor:
  """(b,1,H,W) int64 class ids in [0, C), constant on block×block px tiles."""
  hb, wb = (H + block - 1) // block, (W + block - 1) // block
  ids = (hash_u24(b * hb * wb, seed ^ 0xC1A55, device) % C).reshape(b, hb, wb)
  return ids.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].unsqueeze(1).contiguous()


def block_onehot(b: int, C: int, H: int, W: int, seed: int = 0, block: int = 16,
                 device=None) -> torch.Tensor:
  """(b,C,H,W) float32 one-hot semantics of block_labels (the planes the reference's object-map demo builds from a
  segmentation image, demos/object_map/run.py:117-124)."""
  ids = block_labels(b, C, H, W, seed, block, device)
  out = torch.zeros((b, C, H, W), dtype=torch.float32, device=device)
  out.scatter_(1, ids, 1.0)
  return out


def room_depth(b: int, H: int, W: int, hfov: float, cam_pitch: float, cam_height: float,
               cam_pose: Optional[torch.Tensor] = None, seed: int = 0, half: float = 3.0,
               n_boxes: int = 3, max_depth: float = 10.0, device=None) -> torch.Tensor:
  """(b,1,H,W) depth of an analytic scene ray-cast at each sample's pose: floor plane, four
  walls of a 2*half square room centred on the origin, n_boxes axis-aligned boxes on the floor.
  Coherent depth = realistic same-cell contention.  Depth is the camera-z
  of the hit (what a depth sensor reports), clamped to max_depth."""
  dev = device
  if cam_pose is None:
    cam_pose = torch.zeros((b, 3), dtype=torch.float32, device=dev)
  cam_pose = cam_pose.to(device=dev, dtype=torch.float32).reshape(b, 3)
  cx, cy = W / 2., H / 2.
  fx = cx / math.tan(hfov / 2.)
  fy = fx
  c = torch.arange(W, dtype=torch.float32, device=dev).view(1, 1, W)
  r = torch.arange(H, dtype=torch.float32, device=dev).view(1, H, 1)
  dx = ((c - cx) / fx).expand(b, H, W)
  dy = ((((H - 1) - r) - cy) / fy).expand(b, H, W)
  dz = torch.ones((b, H, W), dtype=torch.float32, device=dev)
  # pitch about x (same sense as the reference: y' = cos*y + sin*z, z' = -sin*y + cos*z)
  cp, sp = math.cos(cam_pitch), math.sin(cam_pitch)
  ly = cp * dy + sp * dz
  lz = -sp * dy + cp * dz
  lx = dx
  yaw = cam_pose[:, 2].view(b, 1, 1)
  cyw, syw = torch.cos(yaw), torch.sin(yaw)
  # yaw about y (reference sense: x' = cos*x - sin*z ... sign irrelevant for a synthetic scene)
  wx = cyw * lx - syw * lz
  wz = syw * lx + cyw * lz
  wy = ly
  ox = cam_pose[:, 0].view(b, 1, 1)
  oz = cam_pose[:, 1].view(b, 1, 1)
  oy = cam_height
  inf = torch.full((b, H, W), float("inf"), dtype=torch.float32, device=dev)

  def plane(o, d, at):
    t = (at - o) / d
    return torch.where((t > 0) & torch.isfinite(t), t, inf)

  t = plane(oy, wy, 0.0)  # floor
  for at in (-half, half):
    t = torch.minimum(t, plane(ox, wx, at))
    t = torch.minimum(t, plane(oz, wz, at))
  # boxes: slab test
  bx = uniform((n_boxes, 4), seed ^ 0xB0C5, 0.0, 1.0, device=dev)
  for k in range(n_boxes):
    cxk = (bx[k, 0] * 2 - 1) * (half - 0.8)
    czk = (bx[k, 1] * 2 - 1) * (half - 0.8)
    hw = 0.2 + 0.4 * bx[k, 2]
    hh = 0.3 + 1.2 * bx[k, 3]
    lo = (cxk - hw, 0.0, czk - hw)
    hi = (cxk + hw, hh, czk + hw)
    tmin = torch.zeros_like(t)
    tmax = inf.clone()
    for o, d, l, h in ((ox, wx, lo[0], hi[0]), (oy, wy, lo[1], hi[1]), (oz, wz, lo[2], hi[2])):
      t1 = (l - o) / d
      t2 = (h - o) / d
      tmin = torch.maximum(tmin, torch.minimum(t1, t2))
      tmax = torch.minimum(tmax, torch.maximum(t1, t2))
    hit = (tmax >= tmin) & (tmin > 0)
    t = torch.where(hit, torch.minimum(t, tmin), t)
  depth = torch.clamp(t, max=max_depth)  # ray has camera-z component 1 → depth == t
  depth = torch.where(torch.isfinite(depth), depth, torch.full_like(depth, max_depth))
  return depth.unsqueeze(1).contiguous()


def frames(kind: str, b: int, H: int, W: int, C: int, seed: int = 0, hfov: float = math.radians(70),
           cam_pitch: float = math.radians(-10), cam_height: float = 0.88, device=None
           ) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
  """(depth (b,1,H,W), values (b,C,H,W) or None, cam_pose (b,3)) for kind in {'iid','room'}."""
  pose = poses(b, seed, device=device)
  if kind == "iid":
    depth = iid_depth(b, H, W, seed, device=device)
  elif kind == "room":
    depth = room_depth(b, H, W, hfov, cam_pitch, cam_height, pose, seed, device=device)
  else:
    raise ValueError(f"unknown synthetic scene kind: {kind}")
  values = block_onehot(b, C, H, W, seed, device=device) if C > 0 else None
  return depth, values, pose
