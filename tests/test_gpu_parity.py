"""GPU parity: the sm_100a kernels (through the public API → C ABI) against
 (a) the golden fixtures produced by the unmodified reference, and
 (b) the pinned CPU oracle on seeded inputs the fixtures do not cover.
Bit-exact for bins / maps / masks / integer outputs; float outputs compare equal value for
value (NaN == NaN, -0 == +0), which is tighter than the 1e-6 / 1e-5 the spec allows.
"""
import math

import numpy as np
import pytest
import torch

import dungeon_maps_b200 as dmap
from dungeon_maps_b200 import synth
from oracle import dm_oracle as orc
from tests._golden import Golden, assert_same, names, sha
from tests.test_oracle_golden import _orth_inputs

pytestmark = pytest.mark.gpu

HFOV = math.radians(70)
PITCH = math.radians(-10)


def npy(t):
  return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _no_device_side_timeouts():
  """After every test: the persistent kernel's dependency waits never timed out (sticky flag,
  word 2 of the workspace control block) and the control block was re-armed."""
  yield
  assert_workspaces_clean()


def assert_workspaces_clean():
  from dungeon_maps_b200 import _native as nat, maps as _maps
  torch.cuda.synchronize()
  nat.device_status()  # raises if a dependency wait timed out (DM_ETIMEOUT through the mapped status word)
  for ws in _maps._workspaces.values():
    ctrl = ws[:16].view(torch.int32).cpu()
    assert int(ctrl[2]) == 0, "a device-side dependency wait timed out"
    assert int(ctrl[0]) == 0 and int(ctrl[3]) == 0, "control block not re-armed"


def run_gpu_orth(g, depth, values, valid):
  kw = g.kwargs
  return dmap.orth_project(
    depth_map=torch.from_numpy(depth), value_map=None if values is None else torch.from_numpy(values),
    valid_map=None if valid is None else torch.from_numpy(valid), cam_pose=g["pose"], width_offset=g["woff"],
    height_offset=g["hoff"], cam_pitch=g["pitch"], cam_height=g["camh"], device="cuda", **kw)


@pytest.mark.parametrize("name", [n for n in names("orth_") if n != "orth_cfg2_iid_b2"])
def test_orth_project_matches_reference(name):
  g = Golden(name)
  depth, values, valid = _orth_inputs(g)
  out = run_gpu_orth(g, depth, values, valid)
  assert out[0].is_cuda and out[1].dtype == torch.bool
  assert_same(npy(out[0]), g["out_topdown"], "topdown")
  assert_same(npy(out[1]), g["out_mask"], "mask")
  if "out_height" in g:
    if values is None:
      assert out[2] is out[0]                      # maps.py:333-334
    else:
      assert out[2].shape == out[0].shape and out[2].stride(1) == 0   # maps.py:349
    assert_same(npy(out[2][:, :1]), g["out_height"], "height")
  else:
    assert len(out) == 2


def test_orth_project_config2_shapes_hashed():
  g = Golden("orth_cfg2_iid_b2")
  depth, values, _ = _orth_inputs(g)
  top, mask, hgt = run_gpu_orth(g, depth, values, None)
  assert sha(npy(top)) == g.meta["sha_topdown"]
  assert sha(npy(mask).astype(np.uint8)) == g.meta["sha_mask"]
  assert_same(npy(hgt[:, :1]), g["out_height"], "height")


def _random_case(seed):
  rng = np.random.default_rng(seed)
  H = int(rng.integers(5, 70)); W = int(rng.integers(5, 90))
  b = int(rng.integers(1, 6))
  C = int(rng.choice([0, 0, 1, 2, 3, 7, 16, 17, 33, 40]))
  Mh = int(rng.integers(3, 80)); Mw = int(rng.integers(3, 80))
  res = float(rng.choice([0.05, 0.1, 0.25, 0.5]))
  intr = orc.intrinsics(W, H, HFOV, None if rng.random() < 0.5 else math.radians(50))
  kw = dict(map_res=res, map_width=Mw, map_height=Mh, focal_x=intr["fx"], focal_y=intr["fy"],
            center_x=intr["cx"], center_y=intr["cy"],
            trunc_depth_min=None if rng.random() < 0.3 else 0.15,
            trunc_depth_max=None if rng.random() < 0.3 else 5.05,
            trunc_height_max=None if rng.random() < 0.6 else 1.2,
            clip_border=None if rng.random() < 0.3 else int(rng.integers(0, 4)),
            to_global=bool(rng.random() < 0.5), flip_h=bool(rng.random() < 0.7),
            fill_value=[None, -np.inf, 0.0, -1.0, 0.5][int(rng.integers(0, 5))],
            reduction=None, get_height_map=bool(rng.random() < 0.7))
  if C > 0 and rng.random() < 0.2:
    kw["reduction"], kw["fill_value"] = "min", [None, np.inf, 0.0][int(rng.integers(0, 3))]
  depth = synth.iid_depth(b, H, W, seed=seed).numpy()
  if rng.random() < 0.3:
    depth.reshape(-1)[::13] = np.nan
  values = None
  if C > 0:
    values = (synth.block_onehot(b, C, H, W, seed=seed, block=3).numpy() if rng.random() < 0.5
              else synth.uniform((b, C, H, W), seed + 7, -2., 2.).numpy())
  valid = (synth.uniform((b, 1, H, W), seed + 9).numpy() > 0.3) if rng.random() < 0.4 else None
  pose = synth.poses(b, seed).numpy()
  woff = (Mw / 2 + rng.normal(size=b)).astype(np.float32)
  hoff = rng.normal(size=b).astype(np.float32) + (Mh / 2 if kw["to_global"] else 0)
  pitch = np.full(b, PITCH, np.float32) + rng.normal(size=b).astype(np.float32) * 0.05
  camh = np.full(b, 0.88, np.float32)
  return depth, values, valid, pose, woff, hoff, pitch, camh, kw


@pytest.mark.parametrize("seed", range(40))
def test_orth_project_random_vs_oracle(seed):
  depth, values, valid, pose, woff, hoff, pitch, camh, kw = _random_case(1000 + seed)
  want = orc.orth_project(depth, values, valid, pose, woff, hoff, pitch, camh, **kw)
  for rep in range(2):  # second call re-uses the (re-zeroed) accumulation ring
    got = dmap.orth_project(torch.from_numpy(depth), None if values is None else torch.from_numpy(values),
                            None if valid is None else torch.from_numpy(valid), pose, woff, hoff, pitch, camh,
                            device="cuda", **kw)
    assert_same(npy(got[0]), want[0], f"topdown rep{rep}")
    assert_same(npy(got[1]), want[1], f"mask rep{rep}")
    if kw["get_height_map"]:
      assert_same(npy(got[2][:, :1]), want[2][:, :1], f"height rep{rep}")


def _tile_case(seed):
  """Shapes the warp-specialised kernel takes (W % 4 == 0), with ragged tile edges in both directions."""
  rng = np.random.default_rng(seed)
  W = 4 * int(rng.integers(1, 80)); H = int(rng.integers(1, 40))
  b = int(rng.integers(1, 14))
  C = int(rng.choice([0, 1, 3, 16, 17, 25, 40, 49]))
  Mh = int(rng.integers(3, 90)); Mw = int(rng.integers(3, 90))
  intr = orc.intrinsics(W, H, HFOV)
  kw = dict(map_res=float(rng.choice([0.05, 0.1, 0.25])), map_width=Mw, map_height=Mh, focal_x=intr["fx"],
            focal_y=intr["fy"], center_x=intr["cx"], center_y=intr["cy"],
            trunc_depth_min=None if rng.random() < 0.3 else 0.15, trunc_depth_max=None if rng.random() < 0.3 else 5.05,
            trunc_height_max=None if rng.random() < 0.6 else 1.2,
            clip_border=None if rng.random() < 0.5 else int(rng.integers(0, 3)),
            to_global=bool(rng.random() < 0.5), flip_h=bool(rng.random() < 0.7),
            fill_value=[None, -np.inf, 0.0, -1.0][int(rng.integers(0, 4))], reduction=None,
            get_height_map=bool(rng.random() < 0.7))
  if C > 0 and rng.random() < 0.25:
    kw["reduction"], kw["fill_value"] = "min", [None, np.inf, 0.0][int(rng.integers(0, 3))]
  # coherent depth (long runs down the columns and along the rows) or i.i.d. depth
  if rng.random() < 0.5:
    depth = synth.room_depth(b, H, W, HFOV, PITCH, 0.88, synth.poses(b, seed), seed).numpy()
  else:
    depth = synth.iid_depth(b, H, W, seed=seed).numpy()
  if rng.random() < 0.3:
    depth.reshape(-1)[::11] = np.nan
  values = None
  if C > 0:
    values = (synth.block_onehot(b, C, H, W, seed=seed, block=3).numpy() if rng.random() < 0.5
              else synth.uniform((b, C, H, W), seed + 7, -2., 2.).numpy())
  valid = (synth.uniform((b, 1, H, W), seed + 9).numpy() > 0.3) if rng.random() < 0.3 else None
  pose = synth.poses(b, seed).numpy()
  woff = (Mw / 2 + rng.normal(size=b)).astype(np.float32)
  hoff = rng.normal(size=b).astype(np.float32) + (Mh / 2 if kw["to_global"] else 0)
  pitch = np.full(b, PITCH, np.float32)
  camh = np.full(b, 0.88, np.float32)
  return depth, values, valid, pose, woff, hoff, pitch, camh, kw


@pytest.mark.parametrize("seed", range(24))
def test_orth_project_tile_layouts_vs_oracle(seed):
  """The float projection kernel has three tile layouts — 2-D tiles of 4 / 8 image rows staged by tensor-map TMA
  copies, tiles of 512 consecutive pixels staged by bulk copies (dm_debug_set_tile_rows) — and every one of them must
  give the oracle's bits."""
  from dungeon_maps_b200 import _native as nat
  depth, values, valid, pose, woff, hoff, pitch, camh, kw = _tile_case(5000 + seed)
  want = orc.orth_project(depth, values, valid, pose, woff, hoff, pitch, camh, **kw)
  try:
    for rows in (4, 8, 0, -1):
      nat.lib().dm_debug_set_tile_rows(rows)
      got = dmap.orth_project(torch.from_numpy(depth), None if values is None else torch.from_numpy(values),
                              None if valid is None else torch.from_numpy(valid), pose, woff, hoff, pitch, camh,
                              device="cuda", **kw)
      assert_same(npy(got[0]), want[0], f"topdown rows={rows}")
      assert_same(npy(got[1]), want[1], f"mask rows={rows}")
      if kw["get_height_map"]:
        assert_same(npy(got[2][:, :1]), want[2][:, :1], f"height rows={rows}")
  finally:
    nat.lib().dm_debug_set_tile_rows(-1)


def test_orth_project_edge_shapes():
  intr = orc.intrinsics(4, 4, HFOV)
  kw = dict(map_res=0.5, map_width=8, map_height=8, focal_x=intr["fx"], focal_y=intr["fy"], center_x=intr["cx"],
            center_y=intr["cy"], trunc_depth_min=None, trunc_depth_max=None, trunc_height_max=None,
            clip_border=None, to_global=False, fill_value=-np.inf, get_height_map=True)
  # single pixel frame, 9*N < 400 → the non-fused rotation path of the reference
  for H, W in ((1, 1), (2, 3), (4, 11), (2, 4), (5, 8), (3, 12)):
    depth = synth.iid_depth(2, H, W, seed=5).numpy()
    k = dict(kw, **{a: v for a, v in zip(("focal_x", "focal_y", "center_x", "center_y"),
                                         (lambda i: (i["fx"], i["fy"], i["cx"], i["cy"]))(orc.intrinsics(W, H, HFOV)))})
    want = orc.orth_project(depth, None, None, np.zeros((2, 3), np.float32), 4., 0., PITCH, 0.88, **k)
    got = dmap.orth_project(torch.from_numpy(depth), None, None, np.zeros((2, 3), np.float32), 4., 0., PITCH, 0.88,
                            device="cuda", **k)
    assert_same(npy(got[0]), want[0], f"topdown {H}x{W}")
    assert_same(npy(got[1]), want[1], f"mask {H}x{W}")
  # everything invalid → all fill, mask all False
  depth = np.full((1, 1, 6, 7), 100.0, np.float32)
  k = dict(kw, trunc_depth_max=5.0, fill_value=0.25)
  top, mask, _ = dmap.orth_project(torch.from_numpy(depth), None, None, [0., 0., 0.], 4., 0., PITCH, 0.88,
                                   device="cuda", **dict(k, focal_x=5., focal_y=5., center_x=3.5, center_y=3.))
  assert (npy(top) == 0.25).all() and not npy(mask).any()
  # an unknown reduction is a ValueError like the reference's Reduction(...) (utils.py:52-67)
  with pytest.raises(ValueError):
    dmap.orth_project(torch.from_numpy(depth), None, None, [0., 0., 0.], 4., 0., PITCH, 0.88, device="cuda",
                      **dict(k, reduction="median", focal_x=5., focal_y=5., center_x=3.5, center_y=3.))


@pytest.mark.parametrize("C,get_height,reduction,to_global", [(40, True, None, False), (33, False, "min", True),
                                                             (64, True, None, True)])
def test_orth_project_many_channels(C, get_height, reduction, to_global):
  """Many value channels (BASELINE config 5 has 40): the kernel switches to 256-pixel tiles with two consumer
  warps per CTA so that enough stages stay resident per SM."""
  b, H, W = 5, 72, 128
  intr = orc.intrinsics(W, H, HFOV)
  kw = dict(map_res=0.05, map_width=120, map_height=100, focal_x=intr["fx"], focal_y=intr["fy"], center_x=intr["cx"],
            center_y=intr["cy"], trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None,
            clip_border=2, to_global=to_global, fill_value=0.25 if reduction else -np.inf, reduction=reduction,
            get_height_map=get_height)
  depth, values, pose = synth.frames("room", b, H, W, C, seed=31)
  values = values * synth.uniform(values.shape, 5, -1.0, 2.0)  # not just {0, 1}
  d, v, p = depth.numpy(), values.numpy(), pose.numpy()
  want = orc.orth_project(d, v, None, p, 60., 0., PITCH, 0.88, **kw)
  for rep in range(2):
    got = dmap.orth_project(depth, values, None, p, 60., 0., PITCH, 0.88, device="cuda", **kw)
    assert_same(npy(got[0]), want[0], f"topdown rep{rep}")
    assert_same(npy(got[1]), want[1], f"mask rep{rep}")
    if get_height:
      assert_same(npy(got[2][:, :1]), want[2][:, :1], f"height rep{rep}")


def test_orth_project_config5_shapes():
  """BASELINE config 5's frame shape (1280x720 depth + 40 semantic channels → 400x400), a few frames bit-exact."""
  b, H, W, C = 3, 720, 1280, 40
  depth, values, pose = synth.frames("room", b, H, W, C, seed=5, device="cuda")
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=200., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.03, map_width=400, map_height=400,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10, to_global=False,
                           fill_value=dmap.NINF)
  top, mask, hgt = proj.orth_project(depth, values, cam_pose=pose, get_height_map=True)
  k = proj.cam_params
  want = orc.orth_project(npy(depth), npy(values), None, npy(pose), 200., 0., PITCH, 0.88, 0.03, 400, 400,
                          k.fx, k.fy, k.cx, k.cy, 0.15, 5.05, None, 10, False, True, -np.inf, None, True, threads=3)
  assert_same(npy(top), want[0], "topdown")
  assert_same(npy(mask), want[1], "mask")
  assert_same(npy(hgt[:, :1]), want[2], "height")


def test_orth_project_depth_channels_fold_into_batch():
  H, W = 20, 24
  intr = orc.intrinsics(W, H, HFOV)
  kw = dict(map_res=0.25, map_width=30, map_height=30, focal_x=intr["fx"], focal_y=intr["fy"], center_x=intr["cx"],
            center_y=intr["cy"], trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None,
            clip_border=1, to_global=False, fill_value=-np.inf, get_height_map=False)
  depth = synth.iid_depth(6, H, W, seed=77).numpy().reshape(2, 3, H, W)
  got = dmap.orth_project(torch.from_numpy(depth), None, None, np.zeros((2, 3), np.float32), 15., 0., PITCH, 0.88,
                          device="cuda", **kw)
  want = orc.orth_project(depth.reshape(6, 1, H, W), None, None, np.zeros((6, 3), np.float32), 15., 0., PITCH, 0.88, **kw)
  # NB: the reference rotates all 3*H*W points of a sample in one bmm; N is large either way
  assert_same(npy(got[0]).reshape(6, 1, 30, 30), want[0], "topdown")


def test_orth_project_full_config2_properties():
  """BASELINE config 2 at full size (64 x 480x640 + 16 channels → 400x400): a slice of frames
  bit-exact against the oracle, plus size-independent properties on the whole batch."""
  b, H, W, C = 64, 480, 640, 16
  depth, values, pose = synth.frames("room", b, H, W, C, seed=3, device="cuda")
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=200., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.03, map_width=400, map_height=400,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10, to_global=False,
                           fill_value=dmap.NINF)
  top, mask, hgt = proj.orth_project(depth, values, cam_pose=pose, get_height_map=True)
  top2, mask2, hgt2 = proj.orth_project(depth, values, cam_pose=pose, get_height_map=True)
  assert torch.equal(top, top2) and torch.equal(mask, mask2) and torch.equal(hgt[:, :1], hgt2[:, :1])  # deterministic
  assert torch.equal(mask, top != dmap.NINF)                       # mask == "cell changed"
  assert torch.equal(mask.any(1, keepdim=True), hgt[:, :1] != dmap.NINF)   # a cell is hit in every channel or none
  vals = top[mask]
  assert ((vals == 0) | (vals == 1)).all()                         # one-hot inputs stay {0,1}
  assert (top.amax(1)[mask.any(1)] == 1).all()                      # every hit cell saw its pixel's class
  sel = [0, 17, 63]
  intr = proj.cam_params
  want = orc.orth_project(npy(depth[sel]), npy(values[sel]), None, npy(pose[sel]), 200., 0., PITCH, 0.88, 0.03, 400, 400,
                          intr.fx, intr.fy, intr.cx, intr.cy, 0.15, 5.05, None, 10, False, True, -np.inf, None, True,
                          threads=3)
  assert_same(npy(top[sel]), want[0], "topdown slice")
  assert_same(npy(mask[sel]), want[1], "mask slice")
  assert_same(npy(hgt[sel][:, :1]), want[2], "height slice")
  assert_workspaces_clean()


def test_orth_project_host_buffer_entry():
  """dm_orth_project_host_f32: HOST buffers in, HOST buffers out (the e2e path of bench.py)."""
  from dungeon_maps_b200 import hostapi
  depth, values, valid, pose, woff, hoff, pitch, camh, kw = _random_case(4242)
  want = orc.orth_project(depth, values, valid, pose, woff, hoff, pitch, camh, **kw)
  got = hostapi.orth_project_host(depth, values, valid, pose, woff, hoff, pitch, camh, **kw)
  assert_same(got[0], want[0], "topdown")
  assert_same(got[1], want[1], "mask")


@pytest.mark.parametrize("chunk", ["1", "2", "3"])
def test_orth_project_host_buffer_pipeline_reuses_slots(chunk):
  """More chunks than staging slots (4): every slot of the three-stream pipeline is re-used, with ragged last
  chunks; results (heights included) must equal the oracle's, call after call."""
  from dungeon_maps_b200 import hostapi
  from dungeon_maps_b200 import _native as nat
  nat.lib().dm_debug_set_host_chunk(int(chunk))
  b, H, W, C = 11, 48, 64, 3
  depth = synth.iid_depth(b, H, W, seed=77).numpy()
  values = synth.uniform((b, C, H, W), 78, -2., 2.).numpy()
  pose = synth.poses(b, 79).numpy()
  intr = orc.intrinsics(W, H, HFOV)
  kw = dict(map_res=0.1, map_width=40, map_height=36, focal_x=intr["fx"], focal_y=intr["fy"], center_x=intr["cx"],
            center_y=intr["cy"], trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None, clip_border=2,
            to_global=False, fill_value=-np.inf, get_height_map=True)
  want = orc.orth_project(depth, values, None, pose, 20., 0., PITCH, 0.88, **kw)
  try:
    for rep in range(2):
      got = hostapi.orth_project_host(depth, values, None, pose, 20., 0., PITCH, 0.88, **kw)
      assert_same(got[0], want[0], f"topdown rep{rep}")
      assert_same(got[1], want[1], f"mask rep{rep}")
      assert_same(got[2][:, :1], want[2][:, :1], f"height rep{rep}")
  finally:
    nat.lib().dm_debug_set_host_chunk(0)


@pytest.mark.parametrize("name", ["flow_small", "flow_small_noflip_vfov"])
def test_camera_affine_grid_matches_reference(name):
  g = Golden(name)
  grid = dmap.camera_affine_grid(torch.from_numpy(g["depth"]), g["pose"], g["pitch"], g["camh"], device="cuda", **g.kwargs)
  assert_same(npy(grid), g["out_grid"], "grid")


def test_camera_affine_grid_480x640_and_flow():
  g = Golden("flow_480x640")
  s = g.meta["synth"]
  depth = synth.iid_depth(s["b"], s["H"], s["W"], seed=s["seed"], device="cuda")
  grid = dmap.camera_affine_grid(depth, g["pose"], PITCH, 0.88, **g.kwargs)
  assert sha(npy(grid)) == g.meta["sha_grid"]
  g2 = Golden("flow_small_egoflow")
  proj = dmap.MapProjector(width=64, height=48, hfov=HFOV, cam_pitch=PITCH, cam_height=0.88)
  flow = dmap.compute_ego_flow(proj, torch.from_numpy(g2["depth"]), g2["pose"])
  assert_same(npy(flow), g2["out_flow"], "ego flow")


def test_camera_affine_grid_host_pipeline():
  """hostapi.camera_affine_grid_host: host depth in, host grid out through the chunked three-stream pipeline —
  ragged last chunk, more chunks than streams, repeated calls — equals the oracle bit for bit."""
  from dungeon_maps_b200 import hostapi
  b, H, W = 11, 48, 64
  depth = synth.iid_depth(b, H, W, seed=31)
  pose = synth.uniform((b, 3), 32, -0.3, 0.3)
  intr = orc.intrinsics(W, H, HFOV)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pitch=PITCH, cam_height=0.88, device="cuda")
  want = orc.camera_affine_grid(depth.numpy(), pose.numpy(), PITCH, 0.88, intr["fx"], intr["fy"], intr["cx"], intr["cy"])
  out = torch.empty((b, 1, H, W, 2), dtype=torch.float32).pin_memory()
  for chunk in (3, 4, 32):
    out.fill_(-7.0)
    got = hostapi.camera_affine_grid_host(proj, depth.pin_memory(), pose, out=out, chunk=chunk)
    assert got is out
    assert_same(got.numpy(), want, f"grid chunk={chunk}")
  assert_same(hostapi.camera_affine_grid_host(proj, depth, pose, chunk=5).numpy(), want, "grid (pageable in, fresh out)")


def test_camera_affine_grid_full_config3():
  """BASELINE config 3 at full size (256 x 480x640, random pose deltas): frames bit-exact against the oracle —
  among them frames salted with NaN / inf / zero / negative / huge depths, which leave the packed straight-line
  path for the generic one inside the same kernel — plus size-independent properties on the whole batch."""
  b, H, W = 256, 480, 640
  depth, _, pose = synth.frames("room", b, H, W, 0, seed=11, device="cuda")
  flat = depth[5].view(-1)
  flat[::17] = float("nan"); flat[5::29] = float("inf"); flat[3::31] = 0.0; flat[7::37] = -1.5
  flat[11::41] = float("-inf"); flat[13::43] = 1e30; flat[1::47] = 1e-30
  delta = (pose * torch.tensor([0.25, 0.25, 0.1], device="cuda")).cpu()
  delta[9] = 0.0                                 # no motion
  delta[10, 2] = 0.0005                          # yaw inside the |a| <= 0.001 clamp (utils.py:323-324)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pitch=PITCH, cam_height=0.88)
  grid = proj.camera_affine_grid(depth, delta)
  assert grid.shape == (b, 1, H, W, 2)
  assert torch.equal(grid.nan_to_num(), proj.camera_affine_grid(depth, delta).nan_to_num())   # deterministic
  sel = [0, 5, 9, 10, 255]
  k = proj.cam_params
  want = orc.camera_affine_grid(npy(depth[sel]), delta[sel].numpy(), PITCH, 0.88, k.fx, k.fy, k.cx, k.cy, threads=5)
  assert_same(npy(grid[sel]), want, "grid slice")
  # no motion: every pixel lands on itself (up to the rounding of the round trip)
  cols = torch.arange(W, dtype=torch.float32, device="cuda").view(1, W)
  rows = torch.arange(H, dtype=torch.float32, device="cuda").view(H, 1)
  assert (grid[9, 0, ..., 0] - cols).abs().max() < 2e-3 and (grid[9, 0, ..., 1] - rows).abs().max() < 2e-3
  # the demo's ego flow is (x - gx, -(y - gy)) of the same grid (demos/ego_flow/run.py:86-89)
  flow = dmap.compute_ego_flow(proj, depth[:1], delta[:1])
  assert torch.equal(flow[..., 0], cols - grid[0, 0, ..., 0]) and torch.equal(flow[..., 1], -(rows - grid[0, 0, ..., 1]))


def test_camera_affine_grid_odd_shapes_vs_oracle():
  for seed, (b, c, H, W) in enumerate([(1, 1, 7, 9), (3, 2, 5, 5), (2, 1, 33, 17), (5, 1, 48, 64)]):
    depth = synth.iid_depth(b * c, H, W, seed=seed).numpy().reshape(b, c, H, W)
    pose = synth.poses(b, seed, xz=0.25, yaw=0.3).numpy()
    intr = orc.intrinsics(W, H, HFOV)
    want = orc.camera_affine_grid(depth, pose, PITCH, 0.88, intr["fx"], intr["fy"], intr["cx"], intr["cy"])
    got = dmap.camera_affine_grid(torch.from_numpy(depth), pose, PITCH, 0.88, intr["fx"], intr["fy"], intr["cx"],
                                  intr["cy"], device="cuda")
    assert_same(npy(got), want, f"grid {b}x{c}x{H}x{W}")


def test_primitives_match_reference():
  g = Golden("primitives")
  cu = lambda a: torch.from_numpy(np.asarray(a)).cuda()
  pose = np.asarray([[0.4, -0.7, 1.1]], np.float32)
  for N in (1, 2, 5, 33, 1000):
    pts, ang = cu(g[f"rot_pts_{N}"]), g[f"rot_ang_{N}"]
    assert_same(npy(dmap.utils.rotate(pts, [1., 0., 0.], ang)), g[f"rot_x_{N}"], f"rotate x {N}")
    assert_same(npy(dmap.utils.rotate(pts, [0., 1., 0.], -ang)), g[f"rot_y_{N}"], f"rotate y {N}")
    assert_same(npy(dmap.utils.rotate(pts, [0.3, -1.2, 0.5], ang)), g[f"rot_axis_{N}"], f"rotate axis {N}")
    assert_same(npy(dmap.utils.translate(pts, [[0.25, -1.5, 3.0]])), g[f"trans_{N}"], f"translate {N}")
    assert_same(npy(dmap.camera_to_local_space(pts, [PITCH], [0.88])), g[f"c2l_{N}"], "c2l")
    assert_same(npy(dmap.local_to_camera_space(pts, [PITCH], [0.88])), g[f"l2c_{N}"], "l2c")
    assert_same(npy(dmap.local_to_global_space(pts, pose)), g[f"l2g_{N}"], "l2g")
    assert_same(npy(dmap.global_to_local_space(pts, pose)), g[f"g2l_{N}"], "g2l")
  for flip in (1, 0):
    xb, zb = dmap.map_quantize(cu(g["q_in_x"]), cu(g["q_in_z"]), [12.5], [-3.25], 0.03, 400, flip_h=bool(flip))
    assert xb.dtype == torch.int64
    assert_same(npy(xb), g[f"q_x_{flip}"], "x_bin")
    assert_same(npy(zb), g[f"q_z_{flip}"], "z_bin")
    x, z = dmap.map_dequantize(xb, zb, [12.5], [-3.25], 0.03, 400, flip_h=bool(flip))
    assert_same(npy(x), g[f"dq_x_{flip}"], "dq x")
    assert_same(npy(z), g[f"dq_z_{flip}"], "dq z")
  k = g.meta["intr_24x32"]
  for flip in (1, 0):
    pts, ok = dmap.depth_map_to_point_cloud(cu(g["d2p_depth"]), cu(g["d2p_valid"]), **k, trunc_depth_min=0.5,
                                            trunc_depth_max=8.0, flip_h=bool(flip))
    assert_same(npy(pts), g[f"d2p_pts_{flip}"], "d2p points")
    assert_same(npy(ok), g[f"d2p_ok_{flip}"], "d2p valid")
    img = dmap.camera_to_image_space(pts, **k, flip_h=bool(flip))
    assert_same(npy(img), g[f"c2i_{flip}"], "camera_to_image")
    assert_same(npy(dmap.image_to_camera_space(img, **k, flip_h=bool(flip))), g[f"i2c_{flip}"], "image_to_camera")
    hm = dmap.height_map_to_point_cloud(cu(g["hm"]), [6.5], [1.0], 0.1, 10, flip_h=bool(flip))
    assert_same(npy(hm), g[f"hm2p_{flip}"], "height_map_to_point_cloud")
  for tag, fill, red in (("ninf_max", -np.inf, None), ("none_max", None, None), ("zero_max", 0., None),
                         ("inf_min", np.inf, "min"), ("none_min", None, "min")):
    cv, m = dmap.project(cu(g["sc_coords"]), cu(g["sc_vals"]), cu(g["sc_valid"]), cu(g["sc_canvas"]),
                         fill_value=fill, reduction=red)
    assert_same(npy(cv), g[f"sc_out_{tag}"], f"project {tag}")
    assert_same(npy(m), g[f"sc_mask_{tag}"], f"project mask {tag}")
  cv, m = dmap.project(cu(g["sc_coords"]), cu(g["sc_vals"]), cu(g["sc_valid"]), cu(g["sc_canvas"]),
                       canvas_masks=cu(g["sc_canvas_masks"]), fill_value=-np.inf)
  assert_same(npy(m), g["sc_mask_or"], "project canvas_masks")
  assert_same(npy(dmap.utils.ravel_index(torch.tensor([[3, 2, 3], [0, 2, 1]]), (6, 5, 4))), g["ravel"], "ravel")


def test_sum_mean_prod_reductions_match_reference():
  """Reduction.sum / mean / prod through project() and orth_project(): the hits of a cell are
  folded in the reference's index order (csrc/dm_ordered.cu), so values AND masks equal the reference's CPU results
  bit for bit (round 1 compared them to 1e-5 and let 1 % of the mask bits differ)."""
  g = Golden("reduce")
  cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
  for tag, fill, red in g.meta["scatter_tags"]:
    cv, m = dmap.project(cu(g["sc_coords"]), cu(g["sc_vals"]), cu(g["sc_valid"]), cu(g["sc_canvas"]),
                         fill_value=fill, reduction=red)
    assert_same(npy(cv), g[f"sc_out_{tag}"], f"project {tag}")
    assert_same(npy(m), g[f"sc_mask_{tag}"], f"project mask {tag}")
  kw = g.kwargs
  for red, fill in (("sum", 0.), ("mean", 0.), ("prod", 1.)):
    top, mask, hgt = dmap.orth_project(cu(g["depth"]), cu(g["values"]), None, g["pose"], 25., 0., PITCH, 0.88,
                                       fill_value=fill, reduction=red, device="cuda", **kw)
    assert_same(npy(top), g[f"orth_top_{red}"], f"orth_project {red} topdown")
    assert_same(npy(mask), g[f"orth_mask_{red}"], f"orth_project {red} mask")
    assert_same(npy(hgt[:, :1]), g[f"orth_height_{red}"], f"orth_project {red} height (max, exact)")
    assert hgt.shape == top.shape and hgt.stride(1) == 0, "height_map is a stride-0 expand (maps.py:349)"


@pytest.mark.parametrize("name", names("builder_"))
def test_map_builder_matches_reference(name):
  g = Golden(name)
  m = g.meta
  H, W = m["H"], m["W"]
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.1, map_width=60, map_height=60,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=3, to_global=m["to_global"],
                           fill_value=m["fill_value"], reduction=m.get("reduction"), device="cuda")
  builder = dmap.MapBuilder(map_projector=proj)
  for t in range(m["steps"]):
    vals = g.get(f"values_{t}")
    local = builder.step(depth_map=g[f"depth_{t}"][0], value_map=None if vals is None else vals[0],
                         cam_pose=g[f"pose_{t}"], center_mode=m["center_mode"], keep_pose=m["keep_pose"])
    assert_same(npy(local.topdown_map), g[f"local_topdown_{t}"], f"step {t} local topdown")
    assert_same(npy(local.mask), g[f"local_mask_{t}"], f"step {t} local mask")
    assert_same(np.asarray(local.proj.width_offset, np.float32).reshape(-1), g[f"local_woff_{t}"].reshape(-1), "local woff")
    wm = builder.world_map
    assert [wm.proj.map_height, wm.proj.map_width] == m["world_shapes"][t], f"step {t} world shape"
    assert_same(np.asarray(wm.proj.width_offset, np.float32).reshape(-1), g[f"world_woff_{t}"].reshape(-1), f"step {t} world woff")
    assert_same(np.asarray(wm.proj.height_offset, np.float32).reshape(-1), g[f"world_hoff_{t}"].reshape(-1), f"step {t} world hoff")
    assert_same(npy(wm.topdown_map), g[f"world_topdown_{t}"], f"step {t} world topdown")
    assert_same(npy(wm.mask), g[f"world_mask_{t}"], f"step {t} world mask")
    assert_same(npy(wm.height_map), g[f"world_height_{t}"], f"step {t} world height")
  assert_same(npy(builder.world_map.get_camera()), g["world_camera"], "get_camera")
  assert_same(npy(builder.world_map.get_origin()), g["world_origin"], "get_origin")


@pytest.mark.parametrize("name", names("canvas_"))
def test_fixed_canvas_builder_matches_reference_ops(name):
  """MapBuilder(fixed_canvas=...) (opt-in in-place merge) against the composition of reference functions
  that defines it (oracle/make_golden.py: canvas_cases)."""
  g = Golden(name)
  m = g.meta
  H, W, C = m["H"], m["W"], m["C"]
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.1, map_width=60, map_height=60,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=3, to_global=True,
                           fill_value=m["fill_value"], device="cuda")
  builder = dmap.MapBuilder(map_projector=proj, fixed_canvas=(m["Hc"], m["Wc"]))
  first = None
  for t in range(m["steps"]):
    vals = g.get(f"values_{t}")
    builder.step(depth_map=g[f"depth_{t}"][0], value_map=None if vals is None else vals[0], cam_pose=g[f"pose_{t}"],
                 to_global=False, width_offset=30., height_offset=0.)
    wm = builder.world_map
    if first is None:
      first = wm.topdown_map.data_ptr()
    assert wm.topdown_map.data_ptr() == first, "the world canvas must be updated in place"
    assert_same(npy(wm.topdown_map), g[f"world_topdown_{t}"], f"step {t} world topdown")
    assert_same(npy(wm.mask), g[f"world_mask_{t}"], f"step {t} world mask")
    if C > 0:
      assert_same(npy(wm.height_map), g[f"world_height_{t}"], f"step {t} world height")
  assert [wm.proj.map_height, wm.proj.map_width] == [m["Hc"], m["Wc"]]


def test_fixed_canvas_builder_random_vs_oracle():
  """Batch of 3 environments, min reduction off, odd canvas, points falling outside the canvas are dropped."""
  b, H, W = 3, 90, 120
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.05, map_width=80, map_height=80,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=2, to_global=True,
                           fill_value=dmap.NINF, device="cuda")
  builder = dmap.MapBuilder(map_projector=proj, fixed_canvas=(101, 97))
  k = orc.intrinsics(W, H, HFOV)
  fx, fy, cx, cy = k["fx"], k["fy"], k["cx"], k["cy"]
  world = None
  pose = torch.zeros(b, 3)
  for t in range(4):
    pose = pose + synth.poses(b, 50 + t, xz=0.8, yaw=0.9)
    depth = synth.room_depth(b, H, W, HFOV, PITCH, 0.88, pose, seed=3)
    builder.step(depth_map=depth, cam_pose=pose, to_global=False, width_offset=40., height_offset=0.)
    top, mask, hgt = orc.orth_project(depth.numpy(), None, None, pose.numpy(), 40., 0., PITCH, 0.88, 0.05, 80, 80,
                                      fx, fy, cx, cy, 0.15, 5.05, None, 2, False, True, -np.inf, None, True)
    src = orc.FuseSource(hgt, mask, None, 40., 0., 0.05, True, False, pose.numpy())
    world = orc.fuse_inplace(world, src, (101, 97), 0.05, True, -np.inf, None)
    assert_same(npy(builder.world_map.topdown_map), world["topdown"], f"step {t} topdown")
    assert_same(npy(builder.world_map.mask), world["mask"], f"step {t} mask")
  assert 0 < int(world["mask"].sum()) < world["mask"].size


def test_map_builder_config4_shapes_vs_oracle():
  """BASELINE config 4's shapes (480x640 depth → 400x400 local height maps at 3 cm, merged into a growing
  global map), 4 environments in one batch (one bounding box over all of them, maps.py:2166-2173), 4 steps: every
  world map bit-exact against the oracle's restatement of fuse_topdown_maps."""
  import sys
  sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
  from bench import BuilderWorkload
  b, H, W, T = 4, 480, 640, 4
  poses = BuilderWorkload.walk(b, T, seed=2, half=4.0)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.03, map_width=400, map_height=400,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10, fill_value=dmap.NINF,
                           to_global=True, device="cuda")
  builder = dmap.MapBuilder(map_projector=proj)
  world = None
  for t in range(T):
    depth = synth.room_depth(b, H, W, HFOV, PITCH, 0.88, poses[t].cuda(), seed=2, half=6.0, device="cuda")
    local = builder.step(depth, cam_pose=poses[t], to_global=False, width_offset=200., height_offset=0.,
                         map_width=400, map_height=400)
    p = poses[t].numpy()
    src = [orc.FuseSource(npy(local.height_map), npy(local.mask), None, 200., 0., 0.03, True, False, p)]
    if world is not None:
      src.insert(0, orc.FuseSource(world["height"], world["mask"], None, world["width_offset"], world["height_offset"],
                                   0.03, True, True, p))
    world = orc.fuse(src, True, p, 0.03, True)
    wm = builder.world_map
    assert [wm.proj.map_height, wm.proj.map_width] == [world["map_height"], world["map_width"]], f"step {t} shape"
    assert_same(np.asarray(wm.proj.width_offset, np.float32).reshape(()), np.float32(world["width_offset"]), "woff")
    assert_same(np.asarray(wm.proj.height_offset, np.float32).reshape(()), np.float32(world["height_offset"]), "hoff")
    assert_same(npy(wm.topdown_map), world["topdown"], f"step {t} world")
    assert_same(npy(wm.mask), world["mask"], f"step {t} world mask")
  assert wm.mask.any(dim=(1, 2, 3)).all()


def test_fuse_tracked_box_equals_a_scan_and_yields_to_edits():
  """fuse_topdown_maps leaves the bounding box of the map it writes for the next merge (dm_fuse_scatter_track_f32 /
  dm_fuse_bbox_seeded_i64).  (1) The tracked box equals what a scan of that map finds; (2) a merge seeded with it
  equals a merge that scans; (3) after an in-place edit of the mask the box is ignored and the result still equals
  the oracle's."""
  b, H, W = 2, 96, 128
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.05, map_width=120, map_height=120,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=2, fill_value=dmap.NINF,
                           to_global=True, device="cuda")
  import sys
  sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
  from bench import BuilderWorkload
  poses = BuilderWorkload.walk(b, 3, seed=5, half=2.0)
  kw = dict(to_global=False, width_offset=60., height_offset=0., map_width=120, map_height=120)
  depth = lambda t: synth.room_depth(b, H, W, HFOV, PITCH, 0.88, poses[t].cuda(), seed=5, half=4.0, device="cuda")
  builder = dmap.MapBuilder(map_projector=proj)
  builder.step(depth(0), cam_pose=poses[0], **kw)
  builder.step(depth(1), cam_pose=poses[1], **kw)
  wm = builder.world_map
  assert getattr(wm, "_tracked_box", None) is not None
  # (1) tracked box == scanned box of the same map
  from dungeon_maps_b200 import maps as _maps, _native as nat
  keep = []
  src = (nat.DmFuseSource * 1)(_maps._fuse_source(wm, wm.proj, b, 1, int(np.prod(wm.mask.shape[1:])), torch.device("cuda", 0), keep))
  scanned = torch.empty((5,), dtype=torch.int64, device="cuda")
  nat.check(nat.lib().dm_fuse_bbox_i64(src, 1, b, 1, wm.proj.map_res, scanned.data_ptr(), nat.stream_ptr(torch.device("cuda", 0))), "bbox")
  assert scanned[:4].tolist() == wm._tracked_box.box[:4].tolist()
  assert int(scanned[4]) == int(wm.mask.sum()) and int(wm._tracked_box.box[4]) > 0
  # (2) seeded merge == scanning merge
  local = builder.plot(depth(2), cam_pose=poses[2], **kw)
  tgt = builder.proj.clone(cam_pose=poses[2])
  seeded = dmap.fuse_topdown_maps(wm, local, map_projector=tgt)
  plain_wm = dmap.TopdownMap(topdown_map=wm.topdown_map, mask=wm.mask, height_map=wm.height_map, map_projector=wm.proj)
  plain = dmap.fuse_topdown_maps(plain_wm, local, map_projector=tgt)
  assert seeded.mask.shape == plain.mask.shape
  assert_same(npy(seeded.topdown_map), npy(plain.topdown_map), "seeded vs scanned topdown")
  assert_same(npy(seeded.mask), npy(plain.mask), "seeded vs scanned mask")
  # (3) the caller marks a far-away cell valid in place: the stale box must not be used
  wm.topdown_map[0, 0, 0, 0] = 1.25
  wm.mask[0, 0, 0, 0] = True
  edited = dmap.fuse_topdown_maps(wm, local, map_projector=tgt)
  p = poses[2].numpy()
  want = orc.fuse([orc.FuseSource(npy(wm.height_map), npy(wm.mask), None, float(wm.proj.width_offset), float(wm.proj.height_offset),
                                  0.05, True, True, p),
                   orc.FuseSource(npy(local.height_map), npy(local.mask), None, 60., 0., 0.05, True, False, p)],
                  True, p, 0.05, True)
  assert [edited.proj.map_height, edited.proj.map_width] == [want["map_height"], want["map_width"]]
  assert_same(npy(edited.topdown_map), want["topdown"], "edited world")
  assert_same(npy(edited.mask), want["mask"], "edited world mask")


def test_crop_matches_reference():
  g = Golden("crop")
  h, w = g.meta["h"], g.meta["w"]
  proj = dmap.MapProjector(width=64, height=48, hfov=HFOV, cam_pose=[0.3, -0.2, 0.4], width_offset=18.5,
                           height_offset=2., cam_pitch=PITCH, cam_height=0.88, map_res=0.1, map_width=w, map_height=h,
                           to_global=True, fill_value=dmap.NINF, device="cuda")
  hm = torch.from_numpy(g["hm"]).cuda()
  mask = torch.from_numpy(g["mask"]).cuda()
  vm = torch.from_numpy(g["vm"]).cuda()
  tm = dmap.TopdownMap(topdown_map=hm, mask=mask, height_map=hm, map_projector=proj)
  tv = dmap.TopdownMap(topdown_map=vm, mask=mask.expand(1, 3, h, w), height_map=hm.expand(1, 3, h, w),
                       map_projector=proj.clone(fill_value=0.))
  for i, (center, cw, ch) in enumerate(g.meta["cases"]):
    c = torch.tensor([center], dtype=torch.int64)
    out = tm.select(c, cw, ch)
    assert out.is_height_map
    assert_same(npy(out.topdown_map), g[f"h{i}_top"], f"crop {i} height")
    assert_same(npy(out.mask), g[f"h{i}_mask"], f"crop {i} mask")
    assert_same(np.asarray(out.proj.width_offset, np.float32).reshape(-1), g[f"h{i}_woff"].reshape(-1), "woff")
    assert_same(np.asarray(out.proj.height_offset, np.float32).reshape(-1), g[f"h{i}_hoff"].reshape(-1), "hoff")
    out = tv.select(c, cw, ch)
    assert_same(npy(out.topdown_map), g[f"v{i}_top"], f"crop {i} values")
    assert_same(npy(out.mask), g[f"v{i}_mask"], f"crop {i} value mask")
    assert_same(npy(out.height_map), g[f"v{i}_height"], f"crop {i} value height")
    out = tv.select(c, cw, ch, fill_value=-7.)
    assert_same(npy(out.topdown_map), g[f"vf{i}_top"], f"crop {i} values fill")


def test_native_library_is_the_path():
  """The kernels really launched from our .so (no silent torch fallback)."""
  from dungeon_maps_b200 import _native as nat
  before = nat.launch_count()
  depth = synth.iid_depth(1, 16, 16, seed=1, device="cuda")
  dmap.camera_affine_grid(depth, [0.1, 0.1, 0.1], PITCH, 0.88, 10., 10., 8., 8.)
  assert nat.launch_count() == before + 1
  maps = open("/proc/self/maps").read()
  assert "libdungeon_maps_b200.so" in maps


def test_parameter_uploads_across_streams_and_many_calls():
  """dm_upload_params recycles a ring of 64 device slots and fences their reuse with markers on the caller's stream:
  more calls than slots, alternating between two streams, every call with poses of its own — each result must be the
  one a fresh single-stream call gives."""
  b, H, W, C = 3, 32, 40, 2
  depth, values, _ = synth.frames("room", b, H, W, C, seed=3, device="cuda")
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=25., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.1, map_width=50, map_height=50,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, to_global=True, device="cuda")
  poses = [synth.poses(b, 1000 + i) for i in range(150)]
  want = []
  for p in poses:
    top, mask, _ = proj.orth_project(depth, values, cam_pose=p, get_height_map=True)
    want.append((top.clone(), mask.clone(), proj.camera_affine_grid(depth, p * 0.1).clone()))
  torch.cuda.synchronize()
  streams = [torch.cuda.Stream(), torch.cuda.Stream()]
  got = []
  for i, p in enumerate(poses):
    with torch.cuda.stream(streams[i % 2]):
      top, mask, _ = proj.orth_project(depth, values, cam_pose=p, get_height_map=True)
      got.append((top, mask, proj.camera_affine_grid(depth, p * 0.1)))
  torch.cuda.synchronize()
  for i, (w, g) in enumerate(zip(want, got)):
    assert torch.equal(w[0], g[0]) and torch.equal(w[1], g[1]), f"call {i}: maps differ"
    assert torch.equal(w[2].nan_to_num(), g[2].nan_to_num()), f"call {i}: grids differ"
