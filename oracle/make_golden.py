"""ORACLE — test infrastructure only.  Generates tests/golden/*.npz by running the
UNMODIFIED reference (oracle/ref_shim.py) on CPU.  Run in the authoring
container:  python -m oracle.make_golden

Each fixture stores the inputs (or the synth seed that regenerates them
exactly), the keyword arguments as JSON, and the reference outputs.  Big dense
outputs are stored as a sha256 over the raw bytes plus a strided sample.
"""
import hashlib
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference, per_sample  # noqa: E402
from dungeon_maps_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
dm = load_reference()
torch.set_num_threads(8)

HFOV = math.radians(70)
PITCH = math.radians(-10)


def sha(a: np.ndarray) -> str:
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, meta, **arrays):
  path = os.path.join(OUT, name + ".npz")
  arrays = {k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items() if v is not None}
  np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
  print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def intr(W, H, hfov=HFOV, vfov=None):
  p = dm.utils.get_camera_intrinsics(W, H, hfov, vfov)
  return dict(focal_x=p.fx, focal_y=p.fy, center_x=p.cx, center_y=p.cy)


# ---------------------------------------------------------------- orth_project

def run_orth(name, depth, values, valid, pose, woff, hoff, pitch, camh, kw, store_inputs=True,
             seed_meta=None):
  b = depth.shape[0]
  batched = dict(depth_map=depth, value_map=values, valid_map=valid, cam_pose=pose,
                 width_offset=woff, height_offset=hoff, cam_pitch=pitch, cam_height=camh)
  out = per_sample(dm.orth_project, b, batched, kw)
  meta = dict(kind="orth_project", kwargs={k: (None if v is None else (v.value if hasattr(v, "value") else v)) for k, v in kw.items()},
              synth=seed_meta)
  arrays = dict(pose=pose, woff=woff, hoff=hoff, pitch=pitch, camh=camh,
                out_topdown=out[0], out_mask=out[1])
  if len(out) > 2:
    # height map comes back as a stride-0 expand over channels: keep one channel
    arrays["out_height"] = out[2][:, :1].contiguous()
    meta["height_is_topdown"] = bool(values is None)
  if store_inputs:
    arrays.update(depth=depth, values=values, valid=valid)
  save(name, meta, **arrays)


def orth_cases():
  # o1: BASELINE config 1 — 480x640, batch 1, height map only, reference defaults of SURVEY §8d
  H, W = 480, 640
  kw = dict(map_res=0.03, map_width=400, map_height=400, **intr(W, H), trunc_depth_min=0.15,
            trunc_depth_max=5.05, trunc_height_max=None, clip_border=10, to_global=False, flip_h=True,
            fill_value=dm.NINF, reduction=None, get_height_map=True)
  depth = synth.iid_depth(1, H, W, seed=0)
  run_orth("orth_cfg1_iid", depth, None, None, torch.zeros(1, 3), torch.tensor([200.]), torch.tensor([0.]),
           torch.tensor([PITCH]), torch.tensor([0.88]), kw, store_inputs=False,
           seed_meta=dict(fn="iid_depth", b=1, H=H, W=W, seed=0))

  # small frames
  H, W = 48, 64
  base = dict(map_res=0.25, map_width=40, map_height=40, **intr(W, H), trunc_depth_min=0.15,
              trunc_depth_max=5.05, trunc_height_max=None, clip_border=2, to_global=False, flip_h=True,
              fill_value=dm.NINF, reduction=None, get_height_map=True)

  def small(b, seed):
    return (synth.iid_depth(b, H, W, seed=seed), synth.poses(b, seed), torch.full((b,), 20.),
            torch.zeros(b), torch.full((b,), PITCH), torch.full((b,), 0.88))

  d, p, wo, ho, pi, ch = small(1, 1)
  run_orth("orth_small_height_local", d, None, None, p, wo, ho, pi, ch, base)

  d, p, wo, ho, pi, ch = small(3, 2)
  vals = synth.uniform((3, 3, H, W), 22, -2.0, 2.0)
  run_orth("orth_small_c3_global", d, vals, None, p, wo + torch.tensor([0., 1.5, -3.25]), ho + 5., pi, ch,
           dict(base, to_global=True))

  d, p, wo, ho, pi, ch = small(2, 3)
  vals = synth.block_onehot(2, 16, H, W, seed=3, block=4)
  valid = synth.uniform((2, 1, H, W), 33) > 0.3
  run_orth("orth_small_onehot16_fill0_valid", d, vals, valid, p, wo, ho, pi, ch, dict(base, fill_value=0.))

  d, p, wo, ho, pi, ch = small(2, 4)
  vals = synth.uniform((2, 4, H, W), 44, -1.0, 3.0)
  run_orth("orth_small_c4_fillnone_noflip", d, vals, None, p, wo, ho + 2., pi, ch,
           dict(base, fill_value=None, trunc_depth_min=None, trunc_depth_max=None, trunc_height_max=1.0,
                clip_border=None, flip_h=False, to_global=True))

  d, p, wo, ho, pi, ch = small(1, 5)
  vals = synth.uniform((1, 2, H, W), 55, -1.0, 1.0)
  run_orth("orth_small_c2_min", d, vals, None, p, wo, ho, pi, ch,
           dict(base, fill_value=float("inf"), reduction=dm.utils.Reduction.min, get_height_map=True))

  d, p, wo, ho, pi, ch = small(1, 6)
  run_orth("orth_small_no_height_out", d, None, None, p, wo, ho, pi, ch, dict(base, get_height_map=False, fill_value=0.))

  # special values in depth, no truncation: NaN/inf/0/negative must drop out via the int64 cast
  d, p, wo, ho, pi, ch = small(1, 7)
  flat = d.view(-1)
  flat[::17] = float("nan"); flat[5::29] = float("inf"); flat[3::31] = 0.0; flat[7::37] = -1.5
  flat[11::41] = float("-inf"); flat[13::43] = 1e30
  run_orth("orth_small_specials", d, None, None, p, wo, ho, pi, ch,
           dict(base, trunc_depth_min=None, trunc_depth_max=None, clip_border=0))

  # angles inside the |a| <= 0.001 clamp (utils.py:323-324)
  d, p, wo, ho, pi, ch = small(2, 8)
  p[:, 2] = torch.tensor([0.0009, -0.001])
  run_orth("orth_small_tiny_angles", d, None, None, p, wo, ho, torch.tensor([0.0005, -0.00099]), ch,
           dict(base, to_global=True))

  # vfov given, non-square pixels, fractional offsets
  d, p, wo, ho, pi, ch = small(1, 9)
  run_orth("orth_small_vfov", d, None, None, p, wo + 0.37, ho - 0.61, pi, ch,
           dict(base, **intr(W, H, HFOV, math.radians(50))))

  # coherent scenes
  H, W = 120, 160
  kw = dict(map_res=0.05, map_width=200, map_height=200, **intr(W, H), trunc_depth_min=0.15,
            trunc_depth_max=5.05, trunc_height_max=None, clip_border=5, to_global=False, flip_h=True,
            fill_value=dm.NINF, reduction=None, get_height_map=True)
  pose = synth.poses(2, 10)
  d = synth.room_depth(2, H, W, HFOV, PITCH, 0.88, pose, seed=10)
  run_orth("orth_room_height_local", d, None, None, pose, torch.full((2,), 100.), torch.zeros(2),
           torch.full((2,), PITCH), torch.full((2,), 0.88), kw)
  vals = synth.block_onehot(2, 16, H, W, seed=11)
  run_orth("orth_room_onehot16_global", d, vals, None, pose, torch.full((2,), 100.), torch.full((2,), 100.),
           torch.full((2,), PITCH), torch.full((2,), 0.88), dict(kw, to_global=True, fill_value=0.))

  # BASELINE config 2 shapes, two frames: 480x640 + 16 one-hot channels → 400x400
  H, W = 480, 640
  kw = dict(map_res=0.03, map_width=400, map_height=400, **intr(W, H), trunc_depth_min=0.15,
            trunc_depth_max=5.05, trunc_height_max=None, clip_border=10, to_global=False, flip_h=True,
            fill_value=dm.NINF, reduction=None, get_height_map=True)
  b = 2
  depth, vals, pose = synth.frames("iid", b, H, W, 16, seed=20)
  batched = dict(depth_map=depth, value_map=vals, valid_map=None, cam_pose=pose,
                 width_offset=torch.full((b,), 200.), height_offset=torch.zeros(b),
                 cam_pitch=torch.full((b,), PITCH), cam_height=torch.full((b,), 0.88))
  top, mask, hgt = per_sample(dm.orth_project, b, batched, kw)
  save("orth_cfg2_iid_b2", dict(kind="orth_project_hashed", kwargs={k: v for k, v in kw.items() if k != "reduction"},
                                synth=dict(fn="frames", kind="iid", b=b, H=H, W=W, C=16, seed=20),
                                sha_topdown=sha(top.numpy()), sha_mask=sha(mask.numpy().astype(np.uint8)),
                                sha_height=sha(hgt[:, :1].contiguous().numpy())),
       mask_packed=np.packbits(mask.numpy()), out_height=hgt[:, :1].contiguous(),
       woff=batched["width_offset"], hoff=batched["height_offset"], pitch=batched["cam_pitch"],
       camh=batched["cam_height"], pose=pose)


# ---------------------------------------------------------- camera_affine_grid

def flow_cases():
  H, W = 48, 64
  b = 3
  depth = synth.iid_depth(b, H, W, seed=100)
  dpose = synth.poses(b, 100, xz=0.25, yaw=0.3)
  kw = dict(**intr(W, H), flip_h=True)
  batched = dict(depth_map=depth, trans_pose=dpose, cam_pitch=torch.full((b,), PITCH), cam_height=torch.full((b,), 0.88))
  grid = per_sample(dm.camera_affine_grid, b, batched, kw)
  save("flow_small", dict(kind="camera_affine_grid", kwargs=kw), depth=depth, pose=dpose,
       pitch=batched["cam_pitch"], camh=batched["cam_height"], out_grid=grid)
  kw2 = dict(**intr(W, H, HFOV, math.radians(55)), flip_h=False)
  grid = per_sample(dm.camera_affine_grid, b, batched, kw2)
  save("flow_small_noflip_vfov", dict(kind="camera_affine_grid", kwargs=kw2), depth=depth, pose=dpose,
       pitch=batched["cam_pitch"], camh=batched["cam_height"], out_grid=grid)
  # demo helper compute_ego_flow (demos/ego_flow/run.py:75-90) on the first sample
  g0 = dm.camera_affine_grid(depth_map=depth[0], trans_pose=dpose[0], cam_pitch=PITCH, cam_height=0.88, **kw)
  x, y = dm.utils.generate_image_coords(depth[0].shape, dtype=torch.float32)
  coords = torch.stack((x, y), dim=-1)
  flow = coords - g0
  flow[..., 0] /= g0.shape[1]
  flow[..., 1] /= g0.shape[0]
  flow[..., 1] = -flow[..., 1]
  save("flow_small_egoflow", dict(kind="compute_ego_flow", kwargs=kw), depth=depth[:1], pose=dpose[:1],
       out_flow=flow[0, 0])
  # full 480x640 frame, hashed + strided sample
  H, W = 480, 640
  depth = synth.iid_depth(1, H, W, seed=101)
  dpose = synth.poses(1, 101, xz=0.25, yaw=0.3)
  kw = dict(**intr(W, H), flip_h=True)
  grid = dm.camera_affine_grid(depth_map=depth, trans_pose=dpose, cam_pitch=torch.tensor([PITCH]),
                               cam_height=torch.tensor([0.88]), **kw)
  save("flow_480x640", dict(kind="camera_affine_grid_hashed", kwargs=kw,
                            synth=dict(fn="iid_depth", b=1, H=H, W=W, seed=101), sha_grid=sha(grid.numpy())),
       pose=dpose, out_grid_sample=grid[:, :, ::7, ::5].contiguous())


# ------------------------------------------------------------------ MapBuilder

def builder_cases():
  H, W = 60, 80

  def run(name, to_global, C, center_mode, steps, keep_pose=False, fill_value=dm.NINF, reduction=None):
    proj = dm.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.1, map_width=60, map_height=60,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=3, to_global=to_global,
                           fill_value=fill_value, reduction=reduction)
    builder = dm.MapBuilder(map_projector=proj)
    pose = torch.zeros(1, 3)
    arrays = {}
    shapes = []
    for t in range(steps):
      step = synth.poses(1, 1000 + t, xz=0.4, yaw=0.5)
      pose = pose + step
      depth = synth.room_depth(1, H, W, HFOV, PITCH, 0.88, pose, seed=7)
      vals = synth.block_onehot(1, C, H, W, seed=200 + t, block=8) if C > 0 else None
      if reduction is not None and C > 0:  # order-dependent reductions: values whose sums / products round
        vals = synth.uniform((1, C, H, W), 200 + t, 0.25, 1.75)
      cam_pose = pose[0].numpy().copy()
      local = builder.step(depth_map=depth[0].numpy(), value_map=None if vals is None else vals[0].numpy(),
                           cam_pose=cam_pose, center_mode=center_mode, keep_pose=keep_pose)
      wm = builder.world_map
      arrays[f"depth_{t}"] = depth
      if vals is not None:
        arrays[f"values_{t}"] = vals
      arrays[f"pose_{t}"] = cam_pose
      arrays[f"local_topdown_{t}"] = local.topdown_map
      arrays[f"local_mask_{t}"] = local.mask
      arrays[f"local_height_{t}"] = local.height_map[:, :1].contiguous() if C > 0 else local.height_map
      arrays[f"local_woff_{t}"] = np.asarray(local.proj.width_offset, dtype=np.float32)
      arrays[f"local_hoff_{t}"] = np.asarray(local.proj.height_offset, dtype=np.float32)
      arrays[f"world_topdown_{t}"] = wm.topdown_map
      arrays[f"world_mask_{t}"] = wm.mask
      arrays[f"world_height_{t}"] = wm.height_map.contiguous()
      arrays[f"world_woff_{t}"] = np.asarray(wm.proj.width_offset, dtype=np.float32)
      arrays[f"world_hoff_{t}"] = np.asarray(wm.proj.height_offset, dtype=np.float32)
      shapes.append([int(wm.proj.map_height), int(wm.proj.map_width)])
      if t == steps - 1:
        arrays["world_camera"] = wm.get_camera()
        arrays["world_origin"] = wm.get_origin()
    save(name, dict(kind="map_builder", to_global=to_global, C=C, center_mode=center_mode, steps=steps,
                    keep_pose=keep_pose, fill_value=None if fill_value is None else float(fill_value),
                    reduction=reduction,
                    world_shapes=shapes, H=H, W=W), **arrays)

  run("builder_global_height", True, 0, "none", 4)
  run("builder_local_height_camera", False, 0, "camera", 4)
  run("builder_global_values_origin", True, 3, "origin", 3, fill_value=0.)
  run("builder_local_values_keep_pose", False, 2, "none", 3, keep_pose=True)
  # Reduction.sum / mean / prod through plot AND merge (MapProjector.project falls back to the projector's reduction,
  # maps.py:1720): order-dependent float results, pinned bit for bit
  run("builder_global_values_sum", True, 2, "none", 3, fill_value=0., reduction="sum")
  run("builder_local_values_mean", False, 2, "none", 3, fill_value=0., reduction="mean")
  run("builder_global_height_prod", True, 0, "none", 3, fill_value=1., reduction="prod")


# ------------------------------------------------------------------ fixed-canvas merge

def canvas_cases():
  """The opt-in fixed-canvas merge is not a reference function; its definition is a composition of
  reference functions, which is what runs here: MapBuilder.plot (local map) → _flattened_topdown_map
  (cells → global points, maps.py:2039-2069) → MapProjector.map_quantize with the canvas' fixed offsets
  (maps.py:944-1019) → project(canvas=, canvas_masks=) (maps.py:1089-1173)."""
  H, W = 60, 80
  Hc, Wc = 150, 170

  def run(name, C, steps, fill_value=dm.NINF):
    proj = dm.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.1, map_width=60, map_height=60,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=3, to_global=True,
                           fill_value=fill_value)
    builder = dm.MapBuilder(map_projector=proj)
    world_proj = proj.clone(to_global=True, width_offset=Wc / 2., height_offset=Hc / 2., map_width=Wc, map_height=Hc)
    Cv = max(C, 1)
    canvas = torch.full((1, Cv, Hc, Wc), float(fill_value))
    cmask = torch.zeros((1, Cv, Hc, Wc), dtype=torch.bool)
    hcanvas = torch.full((1, Cv, Hc, Wc), -np.inf)
    pose = torch.zeros(1, 3)
    arrays = {}
    for t in range(steps):
      pose = pose + synth.poses(1, 2000 + t, xz=0.6, yaw=0.7)
      depth = synth.room_depth(1, H, W, HFOV, PITCH, 0.88, pose, seed=9)
      vals = synth.block_onehot(1, C, H, W, seed=300 + t, block=8) if C > 0 else None
      cam_pose = pose[0].numpy().copy()
      local = builder.plot(depth_map=depth[0].numpy(), value_map=None if vals is None else vals[0].numpy(),
                           cam_pose=cam_pose, to_global=False, width_offset=30., height_offset=0.)
      points, mask, values = dm.maps._flattened_topdown_map(local)
      x_bin, z_bin = world_proj.map_quantize(points[..., 0], points[..., 2])
      coords = torch.stack((z_bin, x_bin), dim=-1)  # as orth_project does, maps.py:318
      if values is None:
        values = points[..., 1]
      # the module-level function with _validate_args=False, as orth_project calls it (maps.py:320-331): the
      # validating path broadcasts the canvas over the point dimension, and MapProjector.project would re-fill
      canvas, cmask = dm.maps.project(coords=coords.clone(), values=values, masks=mask, canvas=canvas,
                                      canvas_masks=cmask, fill_value=None, _validate_args=False)
      if C > 0:
        hcanvas, _ = dm.maps.project(coords=coords.clone(), values=points[..., 1], masks=mask, canvas=hcanvas,
                                     fill_value=None, _validate_args=False)
      arrays[f"depth_{t}"] = depth
      if vals is not None:
        arrays[f"values_{t}"] = vals
      arrays[f"pose_{t}"] = cam_pose
      arrays[f"world_topdown_{t}"] = canvas.clone()
      arrays[f"world_mask_{t}"] = cmask.clone()
      if C > 0:
        arrays[f"world_height_{t}"] = hcanvas.clone()
    save(name, dict(kind="fixed_canvas", C=C, steps=steps, fill_value=float(fill_value), H=H, W=W, Hc=Hc, Wc=Wc),
         **arrays)

  run("canvas_height", 0, 4)
  run("canvas_values_fill0", 3, 3, fill_value=0.)


# ------------------------------------------------------------------ sum / mean / prod reductions

def reduce_cases():
  """Reduction.sum / mean / prod (utils.py:70-76) through project() and orth_project().  sum and prod run on the
  shim's scatter_reduce_ (exact semantics of torch_scatter with out=); mean on the restated scatter_mean."""
  arrays = {}
  N, Mh, Mw = 500, 9, 11
  vals = synth.uniform((2, 3, N), 405, -3., 3.)
  rows = (synth.hash_u24(2 * N, 406) % (Mh + 4)).reshape(2, 1, N) - 2
  cols = (synth.hash_u24(2 * N, 407) % (Mw + 4)).reshape(2, 1, N) - 2
  coords = torch.stack((rows, cols), -1)
  mk = synth.uniform((2, 1, N), 408) > 0.25
  canvas0 = synth.uniform((2, 3, Mh, Mw), 409, -1., 1.)
  arrays.update(sc_vals=vals, sc_coords=coords, sc_valid=mk, sc_canvas=canvas0)
  tags = (("zero_sum", 0., "sum"), ("none_sum", None, "sum"), ("one_prod", 1., "prod"), ("none_prod", None, "prod"),
          ("zero_mean", 0., "mean"), ("none_mean", None, "mean"))
  for tag, fill, red in tags:
    cv, m = dm.project(coords=coords.clone(), values=vals, masks=mk, canvas=canvas0.clone(), fill_value=fill,
                       reduction=red, _validate_args=False)
    arrays[f"sc_out_{tag}"], arrays[f"sc_mask_{tag}"] = cv, m
  H, W = 48, 64
  depth = synth.room_depth(1, H, W, HFOV, PITCH, 0.88, synth.poses(1, 61), seed=61)
  values = synth.uniform((1, 2, H, W), 62, 0.5, 1.5)
  pose = synth.poses(1, 61)
  kw = dict(map_res=0.1, map_width=50, map_height=50, trunc_depth_min=0.15, trunc_depth_max=5.05,
            trunc_height_max=None, clip_border=2, to_global=False, flip_h=True, get_height_map=True, **intr(W, H))
  arrays.update(depth=depth, values=values, pose=pose)
  for red, fill in (("sum", 0.), ("mean", 0.), ("prod", 1.)):
    out = dm.orth_project(depth_map=depth, value_map=values, valid_map=None, cam_pose=pose, width_offset=torch.tensor([25.]),
                          height_offset=torch.tensor([0.]), cam_pitch=torch.tensor([PITCH]), cam_height=torch.tensor([0.88]),
                          fill_value=fill, reduction=red, **kw)
    arrays[f"orth_top_{red}"], arrays[f"orth_mask_{red}"], arrays[f"orth_height_{red}"] = out[0], out[1], out[2][:, :1].contiguous()
  save("reduce", dict(kind="reductions", scatter_tags=[list(t) for t in tags], kwargs=kw, H=H, W=W), **arrays)


# ------------------------------------------------------------------ primitives

def primitive_cases():
  arrays = {}
  meta = dict(kind="primitives")
  # utils.rotate / translate on several N (sgemm code paths differ with N)
  for N in (1, 2, 5, 33, 1000):
    pts = synth.uniform((1, N, 3), 300 + N, -5., 5.)
    ang = torch.tensor([0.7 + 0.01 * N])
    arrays[f"rot_pts_{N}"] = pts
    arrays[f"rot_ang_{N}"] = ang
    arrays[f"rot_x_{N}"] = dm.utils.rotate(pts, [1., 0., 0.], ang)
    arrays[f"rot_y_{N}"] = dm.utils.rotate(pts, [0., 1., 0.], -ang)
    arrays[f"rot_axis_{N}"] = dm.utils.rotate(pts, [0.3, -1.2, 0.5], ang)
    off = torch.tensor([[0.25, -1.5, 3.0]])
    arrays[f"trans_{N}"] = dm.utils.translate(pts, off)
    pose = torch.tensor([[0.4, -0.7, 1.1]])
    arrays[f"c2l_{N}"] = dm.camera_to_local_space(pts, torch.tensor([PITCH]), torch.tensor([0.88]))
    arrays[f"l2c_{N}"] = dm.local_to_camera_space(pts, torch.tensor([PITCH]), torch.tensor([0.88]))
    arrays[f"l2g_{N}"] = dm.local_to_global_space(pts, pose)
    arrays[f"g2l_{N}"] = dm.global_to_local_space(pts, pose)
  # quantize / dequantize
  x = synth.uniform((1, 777), 400, -8., 8.)
  z = synth.uniform((1, 777), 401, -8., 8.)
  for flip in (True, False):
    xb, zb = dm.map_quantize(x, z, torch.tensor([12.5]), torch.tensor([-3.25]), 0.03, 400, flip_h=flip)
    arrays[f"q_x_{int(flip)}"], arrays[f"q_z_{int(flip)}"] = xb, zb
    xd, zd = dm.map_dequantize(xb, zb, torch.tensor([12.5]), torch.tensor([-3.25]), 0.03, 400, flip_h=flip)
    arrays[f"dq_x_{int(flip)}"], arrays[f"dq_z_{int(flip)}"] = xd, zd
  arrays["q_in_x"], arrays["q_in_z"] = x, z
  # image <-> camera
  H, W = 24, 32
  k = intr(W, H)
  depth = synth.iid_depth(1, H, W, 402)
  valid = synth.uniform((1, 1, H, W), 403) > 0.2
  for flip in (True, False):
    pc, ok = dm.depth_map_to_point_cloud(depth, valid, **k, trunc_depth_min=0.5, trunc_depth_max=8.0, flip_h=flip)
    arrays[f"d2p_pts_{int(flip)}"], arrays[f"d2p_ok_{int(flip)}"] = pc, ok
    img = dm.camera_to_image_space(pc, **k, flip_h=flip)
    arrays[f"c2i_{int(flip)}"] = img
    arrays[f"i2c_{int(flip)}"] = dm.image_to_camera_space(img, **k, flip_h=flip)
  arrays["d2p_depth"], arrays["d2p_valid"] = depth, valid
  meta["intr_24x32"] = k
  # height_map_to_point_cloud
  hm = synth.uniform((1, 2, 10, 12), 404, -1., 2.)
  arrays["hm"] = hm
  for flip in (True, False):
    arrays[f"hm2p_{int(flip)}"] = dm.height_map_to_point_cloud(hm, torch.tensor([6.5]), torch.tensor([1.0]), 0.1, 10, flip_h=flip)
  # scatter_tensor / project: random coords incl. out of range, canvas kept when fill is None
  N, Mh, Mw = 500, 9, 11
  vals = synth.uniform((2, 3, N), 405, -3., 3.)
  rows = (synth.hash_u24(2 * N, 406) % (Mh + 4)).reshape(2, 1, N) - 2
  cols = (synth.hash_u24(2 * N, 407) % (Mw + 4)).reshape(2, 1, N) - 2
  coords = torch.stack((rows, cols), -1)
  mk = synth.uniform((2, 1, N), 408) > 0.25
  canvas0 = synth.uniform((2, 3, Mh, Mw), 409, -1., 1.)
  arrays.update(sc_vals=vals, sc_coords=coords, sc_valid=mk, sc_canvas=canvas0)
  for tag, fill, red in (("ninf_max", dm.NINF, None), ("none_max", None, None), ("zero_max", 0., None),
                         ("inf_min", float("inf"), "min"), ("none_min", None, "min")):
    cv, m = dm.project(coords=coords.clone(), values=vals, masks=mk, canvas=canvas0.clone(), fill_value=fill,
                       reduction=red, _validate_args=False)
    arrays[f"sc_out_{tag}"], arrays[f"sc_mask_{tag}"] = cv, m
  cm = synth.uniform((2, 3, Mh, Mw), 410) > 0.7
  cv, m = dm.project(coords=coords.clone(), values=vals, masks=mk, canvas=canvas0.clone(), canvas_masks=cm,
                     fill_value=dm.NINF, _validate_args=False)
  arrays["sc_canvas_masks"], arrays["sc_mask_or"] = cm, m
  arrays["ravel"] = dm.utils.ravel_index(torch.tensor([[3, 2, 3], [0, 2, 1]]), (6, 5, 4))
  # compute_center_offsets
  pose = torch.tensor([[0.43, -1.27, 0.9]])
  for mode in ("none", "origin", "camera"):
    for tg in (True, False):
      wo, ho = dm.compute_center_offsets(pose, torch.tensor([1.5]), torch.tensor([-2.0]), 0.03, 400, 400, tg, mode)
      arrays[f"cco_w_{mode}_{int(tg)}"], arrays[f"cco_h_{mode}_{int(tg)}"] = wo, ho
  arrays["cco_pose"] = pose
  save("primitives", meta, **arrays)


def crop_cases():
  arrays = {}
  h, w = 30, 37
  proj = dm.MapProjector(width=64, height=48, hfov=HFOV, cam_pose=[0.3, -0.2, 0.4], width_offset=18.5, height_offset=2.,
                         cam_pitch=PITCH, cam_height=0.88, map_res=0.1, map_width=w, map_height=h, to_global=True,
                         fill_value=dm.NINF)
  hm = synth.uniform((1, 1, h, w), 500, -1., 2.)
  mask = synth.uniform((1, 1, h, w), 501) > 0.4
  hm = torch.where(mask, hm, torch.full_like(hm, dm.NINF))
  tm = dm.TopdownMap(topdown_map=hm, mask=mask, height_map=hm, map_projector=proj)
  vm = synth.uniform((1, 3, h, w), 502, 0., 1.)
  vmask = mask.expand(1, 3, h, w)
  tv = dm.TopdownMap(topdown_map=vm, mask=vmask, height_map=hm.expand(1, 3, h, w), map_projector=proj.clone(fill_value=0.))
  arrays.update(hm=hm, mask=mask, vm=vm)
  cases = [((18, 15), 20, 16), ((0, 0), 12, 12), ((36, 29), 15, 9), ((5, 27), 41, 33), ((18, 14), 37, 30), ((40, -3), 8, 10)]
  for i, (center, cw, ch) in enumerate(cases):
    c = torch.tensor([center], dtype=torch.int64)
    out = tm.select(c.clone(), cw, ch)
    arrays[f"h{i}_top"], arrays[f"h{i}_mask"] = out.topdown_map, out.mask
    arrays[f"h{i}_woff"] = np.asarray(out.proj.width_offset, np.float32)
    arrays[f"h{i}_hoff"] = np.asarray(out.proj.height_offset, np.float32)
    out = tv.select(c.clone(), cw, ch)
    arrays[f"v{i}_top"], arrays[f"v{i}_mask"], arrays[f"v{i}_height"] = out.topdown_map, out.mask, out.height_map
    out = tv.select(c.clone(), cw, ch, fill_value=-7.)
    arrays[f"vf{i}_top"] = out.topdown_map
  save("crop", dict(kind="crop", cases=[[list(c), cw, ch] for c, cw, ch in cases], h=h, w=w), **arrays)


if __name__ == "__main__":
  os.makedirs(OUT, exist_ok=True)
  which = sys.argv[1:] or ["orth", "flow", "builder", "canvas", "reduce", "prim", "crop"]
  if "orth" in which: orth_cases()
  if "flow" in which: flow_cases()
  if "builder" in which: builder_cases()
  if "canvas" in which: canvas_cases()
  if "reduce" in which: reduce_cases()
  if "prim" in which: primitive_cases()
  if "crop" in which: crop_cases()
