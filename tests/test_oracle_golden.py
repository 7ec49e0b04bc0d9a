"""The oracle (oracle/dm_oracle.c) against the reference's own outputs.

tests/golden/*.npz were produced by oracle/make_golden.py from the UNMODIFIED
reference (Ending2015a/dungeon_maps v0.0.3a1, CPU); this pins the oracle
bit-for-bit.  CPU only.
"""
import numpy as np
import pytest
import torch

from dungeon_maps_b200 import synth
from oracle import dm_oracle as orc
from tests._golden import Golden, assert_same, names, sha


def _orth_inputs(g):
  s = g.meta.get("synth")
  if "depth" in g:
    return g["depth"], g.get("values"), g.get("valid")
  if s["fn"] == "iid_depth":
    return synth.iid_depth(s["b"], s["H"], s["W"], seed=s["seed"]).numpy(), None, None
  if s["fn"] == "frames":
    d, v, _ = synth.frames(s["kind"], s["b"], s["H"], s["W"], s["C"], seed=s["seed"])
    return d.numpy(), None if v is None else v.numpy(), None
  raise KeyError(s)


def run_oracle_orth(g, threads=1):
  depth, values, valid = _orth_inputs(g)
  kw = g.kwargs
  return orc.orth_project(depth, values, valid, g["pose"], g["woff"], g["hoff"], g["pitch"], g["camh"],
                          threads=threads, **kw)


@pytest.mark.parametrize("name", [n for n in names("orth_") if n != "orth_cfg2_iid_b2"])
def test_orth_project_matches_reference(name):
  g = Golden(name)
  out = run_oracle_orth(g)
  assert_same(out[0], g["out_topdown"], "topdown")
  assert_same(out[1], g["out_mask"], "mask")
  if "out_height" in g:
    h = out[2] if out[2].shape[1] == 1 else out[2][:, :1]
    assert_same(h, g["out_height"], "height")


def test_orth_project_config2_shapes_hashed():
  g = Golden("orth_cfg2_iid_b2")
  kw = g.kwargs
  depth, values, _ = _orth_inputs(g)
  top, mask, hgt = orc.orth_project(depth, values, None, g["pose"], g["woff"], g["hoff"], g["pitch"],
                                    g["camh"], threads=2, **kw)
  assert sha(top) == g.meta["sha_topdown"]
  assert sha(mask.astype(np.uint8)) == g.meta["sha_mask"]
  assert sha(hgt) == g.meta["sha_height"]
  assert_same(np.packbits(mask), g["mask_packed"], "mask bits")
  assert_same(hgt, g["out_height"], "height")


@pytest.mark.parametrize("name", ["flow_small", "flow_small_noflip_vfov"])
def test_camera_affine_grid_matches_reference(name):
  g = Golden(name)
  grid = orc.camera_affine_grid(g["depth"], g["pose"], g["pitch"], g["camh"], **g.kwargs)
  assert_same(grid, g["out_grid"], "grid")


def test_ego_flow_matches_demo_helper():
  g = Golden("flow_small_egoflow")
  flow = orc.camera_affine_grid(g["depth"], g["pose"], np.float32(np.radians(-10)), 0.88, emit_flow=True, **g.kwargs)
  assert_same(flow[0, 0], g["out_flow"], "flow")


def test_camera_affine_grid_480x640_hashed():
  g = Golden("flow_480x640")
  s = g.meta["synth"]
  depth = synth.iid_depth(s["b"], s["H"], s["W"], seed=s["seed"]).numpy()
  grid = orc.camera_affine_grid(depth, g["pose"], np.float32(np.radians(-10)), 0.88, **g.kwargs)
  assert sha(grid) == g.meta["sha_grid"]
  assert_same(grid[:, :, ::7, ::5], g["out_grid_sample"], "grid sample")


def test_rodrigues_batched_equals_per_sample():
  ang = np.linspace(-3.1, 3.1, 37).astype(np.float32)
  for axis in ([1., 0., 0.], [0., 1., 0.], [0.3, -1.2, 0.5]):
    batched = orc.rodrigues(axis, ang)
    single = np.concatenate([orc.rodrigues(axis, ang[i:i + 1]) for i in range(len(ang))])
    assert_same(batched, single, f"R {axis}")


def test_primitives_match_reference():
  g = Golden("primitives")
  pitch = np.float32(np.radians(-10))
  one = lambda v: np.asarray([v], np.float32)
  pose = np.asarray([[0.4, -0.7, 1.1]], np.float32)
  for N in (1, 2, 5, 33, 1000):
    pts = g[f"rot_pts_{N}"]
    ang = g[f"rot_ang_{N}"]
    z3 = np.zeros((1, 3), np.float32)
    for tag, axis, a in (("x", [1., 0., 0.], ang), ("y", [0., 1., 0.], -ang), ("axis", [0.3, -1.2, 0.5], ang)):
      st = orc.steps(1, orc.rodrigues(axis, a), z3, N)
      # a bare rotate is ROT_THEN_ADD with t = 0 only up to the sign of zero; assert_same ignores it
      assert_same(orc.transform_points(pts, [st]), g[f"rot_{tag}_{N}"], f"rotate {tag} N={N}")
    st = orc.steps(2, orc.rodrigues([1., 0., 0.], one(0.)), np.asarray([[0.25, -1.5, 3.0]], np.float32), N)
    assert_same(orc.transform_points(pts, [st]), g[f"trans_{N}"], f"translate N={N}")
    assert_same(orc.transform_points(pts, [orc.to_local_steps(one(pitch), one(0.88), N)]), g[f"c2l_{N}"], "c2l")
    assert_same(orc.transform_points(pts, [orc.to_camera_steps(one(pitch), one(0.88), N)]), g[f"l2c_{N}"], "l2c")
    assert_same(orc.transform_points(pts, [orc.to_global_steps(pose, N)]), g[f"l2g_{N}"], "l2g")
    assert_same(orc.transform_points(pts, [orc.from_global_steps(pose, N)]), g[f"g2l_{N}"], "g2l")
  for flip in (1, 0):
    xb, zb = orc.map_quantize(g["q_in_x"], g["q_in_z"], 12.5, -3.25, 0.03, 400, flip)
    assert_same(xb, g[f"q_x_{flip}"], "x_bin")
    assert_same(zb, g[f"q_z_{flip}"], "z_bin")
    x, z = orc.map_dequantize(xb, zb, 12.5, -3.25, 0.03, 400, flip)
    assert_same(x, g[f"dq_x_{flip}"], "dq x")
    assert_same(z, g[f"dq_z_{flip}"], "dq z")
  k = g.meta["intr_24x32"]
  for flip in (1, 0):
    pts, ok = orc.depth_to_points(g["d2p_depth"], g["d2p_valid"], k["focal_x"], k["focal_y"], k["center_x"],
                                  k["center_y"], flip, 0.5, 8.0)
    assert_same(pts, g[f"d2p_pts_{flip}"], "d2p points")
    assert_same(ok, g[f"d2p_ok_{flip}"], "d2p valid")
    img = orc.image_camera(pts, k["focal_x"], k["focal_y"], k["center_x"], k["center_y"], flip, 24, 1)
    assert_same(img, g[f"c2i_{flip}"], "camera_to_image")
    cam = orc.image_camera(img, k["focal_x"], k["focal_y"], k["center_x"], k["center_y"], flip, 24, 0)
    assert_same(cam, g[f"i2c_{flip}"], "image_to_camera")
  for tag, fill, red in (("ninf_max", -np.inf, None), ("none_max", None, None), ("zero_max", 0., None),
                         ("inf_min", np.inf, "min"), ("none_min", None, "min")):
    cv, m = orc.scatter(g["sc_vals"], g["sc_coords"], g["sc_valid"], g["sc_canvas"], fill, red)
    assert_same(cv, g[f"sc_out_{tag}"], f"scatter {tag}")
    assert_same(m, g[f"sc_mask_{tag}"], f"scatter mask {tag}")


def test_sum_mean_prod_scatter_matches_reference():
  """Reduction.sum / mean / prod of scatter_tensor (utils.py:70-76, 389-492): the oracle applies the hits in index
  order like the reference's CPU scatter, so even the float sums are bit-identical here."""
  g = Golden("reduce")
  for tag, fill, red in g.meta["scatter_tags"]:
    cv, m = orc.scatter(g["sc_vals"], g["sc_coords"], g["sc_valid"], g["sc_canvas"], fill, red)
    assert_same(cv, g[f"sc_out_{tag}"], f"scatter {tag}")
    assert_same(m, g[f"sc_mask_{tag}"], f"scatter mask {tag}")


def _builder_sources(g, t, C, to_global):
  """(world map after step t-1, local map of step t) as oracle FuseSource objects."""
  srcs = []
  if t > 0:
    wt, wm, wh = g[f"world_topdown_{t-1}"], g[f"world_mask_{t-1}"], g[f"world_height_{t-1}"]
    # keep_pose: the world map stays in the frame of the builder's default pose [0, 0, 0] (maps.py:2496-2497)
    prev_pose = g[f"pose_{t-1}"] if not g.meta["keep_pose"] else np.zeros(3, np.float32)
    srcs.append(orc.FuseSource(wh, wm, wt if C > 0 else None, g[f"world_woff_{t-1}"], g[f"world_hoff_{t-1}"],
                               0.1, True, to_global, prev_pose))
  lt, lm, lh = g[f"local_topdown_{t}"], g[f"local_mask_{t}"], g[f"local_height_{t}"]
  if C > 0:
    lh = np.broadcast_to(lh, lt.shape)
  srcs.append(orc.FuseSource(lh, lm, lt if C > 0 else None, g[f"local_woff_{t}"], g[f"local_hoff_{t}"],
                             0.1, True, to_global, g[f"pose_{t}"]))
  return srcs


@pytest.mark.parametrize("name", names("builder_"))
def test_fuse_topdown_maps_matches_reference(name):
  g = Golden(name)
  C, to_global = g.meta["C"], g.meta["to_global"]
  fill = g.meta["fill_value"]
  for t in range(g.meta["steps"]):
    srcs = _builder_sources(g, t, C, to_global)
    tgt_pose = np.zeros(3, np.float32) if g.meta["keep_pose"] else g[f"pose_{t}"]
    out = orc.fuse(srcs, to_global, tgt_pose, 0.1, True, None, fill, g.meta.get("reduction"))
    assert [out["map_height"], out["map_width"]] == g.meta["world_shapes"][t], f"step {t} shape"
    assert_same(np.float32(out["width_offset"]), g[f"world_woff_{t}"].reshape(()), f"step {t} woff")
    assert_same(np.float32(out["height_offset"]), g[f"world_hoff_{t}"].reshape(()), f"step {t} hoff")
    assert_same(out["topdown"], g[f"world_topdown_{t}"], f"step {t} topdown")
    assert_same(out["mask"], g[f"world_mask_{t}"], f"step {t} mask")
    assert_same(out["height"], g[f"world_height_{t}"], f"step {t} height")


def _canvas_local(g, t, C):
  """The local map of step t through the oracle (the fixture stores inputs and world canvases only)."""
  m = g.meta
  k = orc.intrinsics(m["W"], m["H"], np.radians(70))
  fx, fy, cx, cy = k["fx"], k["fy"], k["cx"], k["cy"]
  vals = g.get(f"values_{t}")
  return orc.orth_project(g[f"depth_{t}"], vals, None, g[f"pose_{t}"], 30., 0., np.radians(-10), 0.88, 0.1, 60, 60,
                          fx, fy, cx, cy, 0.15, 5.05, None, 3, False, True, m["fill_value"], None, True)


@pytest.mark.parametrize("name", names("canvas_"))
def test_fixed_canvas_merge_matches_reference_ops(name):
  """fuse_inplace (restatement of the opt-in fixed-canvas merge) against the composition of reference
  functions that defines it (oracle/make_golden.py: canvas_cases)."""
  g = Golden(name)
  m = g.meta
  C = m["C"]
  world = None
  for t in range(m["steps"]):
    top, mask, hgt = _canvas_local(g, t, C)
    if C > 0:
      hgt = np.broadcast_to(hgt, top.shape)
    src = orc.FuseSource(hgt, mask, top if C > 0 else None, 30., 0., 0.1, True, False, g[f"pose_{t}"])
    world = orc.fuse_inplace(world, src, (m["Hc"], m["Wc"]), 0.1, True, m["fill_value"], None)
    assert_same(world["topdown"], g[f"world_topdown_{t}"], f"step {t} topdown")
    assert_same(world["mask"], g[f"world_mask_{t}"], f"step {t} mask")
    if C > 0:
      assert_same(world["height"], g[f"world_height_{t}"], f"step {t} height")


def test_crop_matches_reference():
  g = Golden("crop")
  for i, (center, cw, ch) in enumerate(g.meta["cases"]):
    top = orc.crop_nearest(g["hm"], center, cw, ch, -np.inf)
    assert_same(top, g[f"h{i}_top"], f"crop {i} height")
    m = orc.crop_nearest(g["mask"].astype(np.float32), center, cw, ch, 0.0) != 0
    assert_same(m, g[f"h{i}_mask"], f"crop {i} mask")
    v = orc.crop_nearest(g["vm"], center, cw, ch, 0.0)
    assert_same(v, g[f"v{i}_top"], f"crop {i} values")
    v = orc.crop_nearest(g["vm"], center, cw, ch, -7.0)
    assert_same(v, g[f"vf{i}_top"], f"crop {i} values fill")
