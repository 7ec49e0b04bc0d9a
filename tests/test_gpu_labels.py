"""GPU parity of the class-id projection (csrc/dm_labels.cu, `orth_project(label_map=, num_classes=)`): bit-identical
to the float path / the reference on value_map = one_hot(label_map).float() — against the reference-generated
fixtures that carry one-hot semantics, against the pinned oracle on seeded cases, and at BASELINE config 2 in full
(all 64 frames hashed).  Plus the delivery of DM_ETIMEOUT and a stress run of the persistent kernels' cross-CTA
protocol (thousands of launches, identical hashes).
"""
import hashlib
import math

import numpy as np
import pytest
import torch

import dungeon_maps_b200 as dmap
from dungeon_maps_b200 import _native as nat
from dungeon_maps_b200 import synth
from oracle import dm_oracle as orc
from tests._golden import Golden, assert_same, names
from tests.test_gpu_parity import assert_workspaces_clean, npy
from tests.test_oracle_golden import _orth_inputs

pytestmark = pytest.mark.gpu

HFOV = math.radians(70)
PITCH = math.radians(-10)


@pytest.fixture(autouse=True)
def _clean():
  yield
  assert_workspaces_clean()


def one_hot(labels, C):
  """(b,1,H,W) integer ids → (b,C,H,W) float32 planes; ids outside [0, C) give an all-zero row."""
  return (np.arange(C).reshape(1, C, 1, 1) == labels.astype(np.int64)).astype(np.float32)


ONEHOT_GOLDENS = [n for n in names("orth_") if "onehot" in n]


@pytest.mark.parametrize("name", ONEHOT_GOLDENS)
def test_labels_match_reference_onehot_fixtures(name):
  """Fixtures the unmodified reference produced from one-hot value maps: the class ids recovered from the planes,
  through the label entry, must give the reference's own outputs."""
  g = Golden(name)
  depth, values, valid = _orth_inputs(g)
  assert values is not None and set(np.unique(values)) <= {0.0, 1.0} and (values.sum(1) == 1).all()
  C = values.shape[1]
  labels = values.argmax(1)[:, None].astype(np.uint8)
  out = dmap.orth_project(
    depth_map=torch.from_numpy(depth), value_map=None, valid_map=None if valid is None else torch.from_numpy(valid),
    cam_pose=g["pose"], width_offset=g["woff"], height_offset=g["hoff"], cam_pitch=g["pitch"], cam_height=g["camh"],
    device="cuda", label_map=torch.from_numpy(labels), num_classes=C, **g.kwargs)
  assert_same(npy(out[0]), g["out_topdown"], "topdown")
  assert_same(npy(out[1]), g["out_mask"], "mask")
  if "out_height" in g:
    assert out[2].shape == out[0].shape and out[2].stride(1) == 0   # maps.py:349
    assert_same(npy(out[2][:, :1]), g["out_height"], "height")


def _random_label_case(seed):
  rng = np.random.default_rng(seed)
  H = int(rng.integers(3, 70)); W = int(rng.integers(3, 90))
  b = int(rng.integers(1, 6))
  C = int(rng.choice([1, 2, 5, 16, 31, 32, 40, 63]))
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
            fill_value=[None, -np.inf, 0.0, -1.0, 0.5, 1.0, 2.0, np.nan][int(rng.integers(0, 8))],
            reduction=None, get_height_map=bool(rng.random() < 0.7))
  if rng.random() < 0.3:
    kw["reduction"], kw["fill_value"] = "min", [None, np.inf, 0.0, 0.5, 1.0, 2.0][int(rng.integers(0, 6))]
  depth = synth.iid_depth(b, H, W, seed=seed, lo=0.1, hi=6.0).numpy()
  if rng.random() < 0.5:   # coherent depth: long same-cell runs that cross pixel quads and lanes
    depth = np.round(depth * 2) / 2 + 0.3
  if rng.random() < 0.3:
    depth.reshape(-1)[::13] = np.nan
  blk = int(rng.choice([1, 3, 8]))
  hi = C + (3 if rng.random() < 0.3 else 0)   # some ids out of range: all-zero one-hot rows
  labels = (synth.hash_u24(b * H * W, seed ^ 0x1AB).numpy().reshape(b, 1, H, W) // 7) % hi
  labels = np.ascontiguousarray(np.repeat(np.repeat(labels[:, :, ::blk, ::blk], blk, 2), blk, 3)[:, :, :H, :W])
  valid = (synth.uniform((b, 1, H, W), seed + 9).numpy() > 0.3) if rng.random() < 0.4 else None
  pose = synth.poses(b, seed).numpy()
  woff = (Mw / 2 + rng.normal(size=b)).astype(np.float32)
  hoff = rng.normal(size=b).astype(np.float32) + (Mh / 2 if kw["to_global"] else 0)
  pitch = np.full(b, PITCH, np.float32) + rng.normal(size=b).astype(np.float32) * 0.05
  camh = np.full(b, 0.88, np.float32)
  return depth, labels, C, valid, pose, woff, hoff, pitch, camh, kw


@pytest.mark.parametrize("seed", range(48))
def test_labels_random_vs_oracle(seed):
  depth, labels, C, valid, pose, woff, hoff, pitch, camh, kw = _random_label_case(7000 + seed)
  want = orc.orth_project(depth, one_hot(labels, C), valid, pose, woff, hoff, pitch, camh, **kw)
  dtype = [np.uint8, np.int64, np.int32][seed % 3]
  lab = labels.astype(dtype)
  for rep in range(2):  # second call re-uses the (re-zeroed) accumulation ring
    got = dmap.orth_project(torch.from_numpy(depth), None, None if valid is None else torch.from_numpy(valid), pose,
                            woff, hoff, pitch, camh, device="cuda", label_map=torch.from_numpy(lab), num_classes=C, **kw)
    assert_same(npy(got[0]), want[0], f"topdown rep{rep}")
    assert_same(npy(got[1]), want[1], f"mask rep{rep}")
    if kw["get_height_map"]:
      assert_same(npy(got[2][:, :1]), want[2][:, :1], f"height rep{rep}")


def test_labels_more_frames_than_ring_slots():
  """b > 64 frames in one call: ring slots are reused inside the launch (projection items then wait for the resolve
  of the slot's previous tenant, resolve completions are published) — b <= 64 gives every frame a slot of its own."""
  b, H, W, C = 150, 24, 32, 7
  depth = synth.iid_depth(b, H, W, seed=77, lo=0.2, hi=4.0).numpy()
  labels = ((synth.hash_u24(b * H * W, 78).numpy().reshape(b, 1, H, W) // 3) % C).astype(np.uint8)
  pose = synth.poses(b, 79).numpy()
  intr = orc.intrinsics(W, H, HFOV)
  kw = dict(map_res=0.1, map_width=40, map_height=36, focal_x=intr["fx"], focal_y=intr["fy"], center_x=intr["cx"],
            center_y=intr["cy"], trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None, clip_border=1,
            to_global=False, fill_value=0.0, get_height_map=True)
  want = orc.orth_project(depth, one_hot(labels, C), None, pose, 20., 0., PITCH, 0.88, **kw)
  for rep in range(2):
    got = dmap.orth_project(torch.from_numpy(depth), None, None, pose, 20., 0., PITCH, 0.88, device="cuda",
                            label_map=torch.from_numpy(labels), num_classes=C, **kw)
    assert_same(npy(got[0]), want[0], f"topdown rep{rep}")
    assert_same(npy(got[1]), want[1], f"mask rep{rep}")
    assert_same(npy(got[2][:, :1]), want[2][:, :1], f"height rep{rep}")


def test_labels_equal_float_path_and_argument_checks():
  """Same call with value_map = one_hot(labels): identical tensors; misuse raises like the rest of the API."""
  b, H, W, C = 3, 60, 80, 16
  depth, values, pose = synth.frames("room", b, H, W, C, seed=2, device="cuda")
  labels = values.argmax(1, keepdim=True).to(torch.uint8)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=50., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.06, map_width=100, map_height=100,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=3, to_global=False, fill_value=0.)
  a = proj.orth_project(depth, values, cam_pose=pose, get_height_map=True)
  l = proj.orth_project(depth, cam_pose=pose, get_height_map=True, label_map=labels, num_classes=C)
  for x, y_, what in zip(a, l, ("topdown", "mask", "height")):
    assert_same(npy(y_), npy(x), what)
  # a shared (1, 1, H, W) label / valid map is broadcast over the batch like any tensor operand of the reference
  shared = proj.orth_project(depth, cam_pose=pose, label_map=labels[:1], num_classes=C,
                             valid_map=torch.ones((H, W), dtype=torch.bool))
  want = proj.orth_project(depth, cam_pose=pose, label_map=labels[:1].expand(b, 1, H, W).contiguous(), num_classes=C)
  assert_same(npy(shared[0]), npy(want[0]), "broadcast labels")
  # the other reductions take the composed path on materialised planes
  s = proj.orth_project(depth, cam_pose=pose, label_map=labels, num_classes=C, reduction="sum")
  s2 = proj.orth_project(depth, values, cam_pose=pose, reduction="sum")
  assert torch.allclose(s[0], s2[0]) and torch.equal(s[1], s2[1])
  with pytest.raises(ValueError):
    proj.orth_project(depth, values, cam_pose=pose, label_map=labels, num_classes=C)
  with pytest.raises(ValueError):
    proj.orth_project(depth, cam_pose=pose, label_map=labels)
  with pytest.raises(ValueError):
    proj.orth_project(depth, cam_pose=pose, label_map=labels, num_classes=64)
  with pytest.raises(TypeError):
    proj.orth_project(depth, cam_pose=pose, label_map=labels.float(), num_classes=C)
  with pytest.raises(RuntimeError):
    proj.orth_project(depth, cam_pose=pose, label_map=labels[:2], num_classes=C)
  with pytest.raises(RuntimeError):  # a value_map whose batch is neither 1 nor b is an error, not an OOB read
    proj.orth_project(depth, values[:2], cam_pose=pose)


def test_builder_with_label_maps():
  """MapBuilder.step(label_map=, num_classes=) == MapBuilder.step(value_map=one_hot): local and world maps."""
  b, H, W, C = 2, 96, 128, 5
  mk = lambda: dmap.MapBuilder(dmap.MapProjector(
    width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=0., height_offset=0., cam_pitch=PITCH,
    cam_height=0.88, map_res=0.05, map_width=120, map_height=120, trunc_depth_min=0.15, trunc_depth_max=5.05,
    clip_border=2, fill_value=0., to_global=True, device="cuda"))
  ba, bl = mk(), mk()
  for t in range(3):
    depth, values, pose = synth.frames("room", b, H, W, C, seed=40 + t, device="cuda")
    labels = values.argmax(1, keepdim=True).to(torch.uint8)
    kw = dict(cam_pose=pose, to_global=False, width_offset=60., height_offset=0., center_mode=dmap.CenterMode.none)
    la = ba.step(depth, value_map=values, **kw)
    ll = bl.step(depth, label_map=labels, num_classes=C, **kw)
    assert not ll.is_height_map
    assert_same(npy(ll.topdown_map), npy(la.topdown_map), f"local map t={t}")
    for x, y_ in ((ba.world_map.topdown_map, bl.world_map.topdown_map), (ba.world_map.mask, bl.world_map.mask),
                  (ba.world_map.height_map, bl.world_map.height_map)):
      assert_same(npy(y_), npy(x), f"world map t={t}")


def _frame_hashes(*arrays):
  """One sha256 per frame over the frame's slices of all arrays."""
  out = []
  for i in range(arrays[0].shape[0]):
    h = hashlib.sha256()
    for a in arrays:
      h.update(np.ascontiguousarray(a[i]).tobytes())
    out.append(h.hexdigest())
  return out


@pytest.mark.parametrize("scene", ["room", "iid"])
def test_full_config2_all_frames_hashed_float_and_labels(scene):
  """BASELINE config 2 at full size, every one of the 64 frames: sha256 per frame of (topdown, mask, height) from
  the float kernel, from the label kernel and from the OpenMP oracle are identical."""
  b, H, W, C = 64, 480, 640, 16
  depth, values, pose = synth.frames(scene, b, H, W, C, seed=3, device="cuda")
  labels = values.argmax(1, keepdim=True).to(torch.uint8)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=200., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.03, map_width=400, map_height=400,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10, to_global=False,
                           fill_value=dmap.NINF)
  k = proj.cam_params
  import os
  want = orc.orth_project(npy(depth), npy(values), None, npy(pose), 200., 0., PITCH, 0.88, 0.03, 400, 400, k.fx, k.fy,
                          k.cx, k.cy, 0.15, 5.05, None, 10, False, True, -np.inf, None, True,
                          threads=min(os.cpu_count() or 1, 32))
  want_h = _frame_hashes(want[0], want[1].astype(np.uint8), want[2])
  for what, kw in (("float", dict(value_map=values)), ("labels", dict(label_map=labels, num_classes=C))):
    top, mask, hgt = proj.orth_project(depth, cam_pose=pose, get_height_map=True, **kw)
    got_h = _frame_hashes(npy(top), npy(mask).astype(np.uint8), npy(hgt[:, :1]))
    bad = [i for i in range(b) if got_h[i] != want_h[i]]
    assert not bad, f"{what} path: frames {bad} differ from the oracle"


def test_labels_host_buffer_entry():
  """dm_orth_project_labels_host_f32 (the e2e path of bench.py): chunked pipeline with slot re-use, ragged tail."""
  from dungeon_maps_b200 import hostapi
  depth, labels, C, valid, pose, woff, hoff, pitch, camh, kw = _random_label_case(9191)
  b = 11
  depth = np.concatenate([depth] * 11)[:b]; labels = np.concatenate([labels] * 11)[:b].astype(np.uint8)
  valid = None if valid is None else np.concatenate([valid] * 11)[:b]
  rep = lambda a: np.resize(a, (b,) + a.shape[1:])
  pose, woff, hoff, pitch, camh = rep(pose), rep(woff), rep(hoff), rep(pitch), rep(camh)
  want = orc.orth_project(depth, one_hot(labels, C), valid, pose, woff, hoff, pitch, camh, **kw)
  for chunk in (0, 2, 3):
    nat.lib().dm_debug_set_host_chunk(chunk)
    try:
      got = hostapi.orth_project_host(depth, None, valid, pose, woff, hoff, pitch, camh, label_map=labels,
                                      num_classes=C, **kw)
    finally:
      nat.lib().dm_debug_set_host_chunk(0)
    assert_same(got[0], want[0], f"topdown chunk={chunk}")
    assert_same(got[1], want[1], f"mask chunk={chunk}")
    if kw["get_height_map"]:
      assert_same(got[2][:, :1], want[2][:, :1], f"height chunk={chunk}")


# ---- DM_ETIMEOUT is delivered, and the workspace survives it ---------------------------------------------------

@pytest.mark.parametrize("path", ["float", "labels"])
def test_device_side_timeout_is_reported_and_recovered(path):
  """With the test hook every cross-CTA dependency is made unsatisfiable and the guard time 0.2 ms: the launch
  ends (no hang), the NEXT entry on the device returns DM_ETIMEOUT once, and the call after that — on the same
  workspace, which the timed-out launch re-zeroed itself — is bit-exact again."""
  b, H, W, C = 24, 60, 80, 4
  depth, values, pose = synth.frames("room", b, H, W, C, seed=8, device="cuda")
  labels = values.argmax(1, keepdim=True).to(torch.uint8)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=50., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.06, map_width=100, map_height=100,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=3, to_global=False)
  kw = dict(value_map=values) if path == "float" else dict(label_map=labels, num_classes=C)
  call = lambda: proj.orth_project(depth, cam_pose=pose, get_height_map=True, **kw)
  good = call()
  torch.cuda.synchronize()
  lib = nat.lib()
  lib.dm_debug_set_wait_guard(200_000, 1)
  try:
    call()                       # every dependent item times out and is skipped: garbage outputs
    torch.cuda.synchronize()
  finally:
    lib.dm_debug_set_wait_guard(0, 0)
  with pytest.raises(nat.NativeError, match="timed out"):
    call()                       # reported once, nothing launched
  again = call()
  torch.cuda.synchronize()
  nat.device_status()            # the flag was consumed
  for x, y_, what in zip(good, again, ("topdown", "mask", "height")):
    assert_same(npy(y_), npy(x), f"{what} after recovery")


def test_timeout_through_the_host_entry():
  from dungeon_maps_b200 import hostapi
  depth, labels, C, valid, pose, woff, hoff, pitch, camh, kw = _random_label_case(9292)
  b = 40
  tile = lambda a: np.resize(a, (b,) + a.shape[1:])
  args = (tile(depth), None, None, tile(pose), tile(woff), tile(hoff), tile(pitch), tile(camh))
  lk = dict(label_map=tile(labels.astype(np.uint8)), num_classes=C, **kw)
  good = hostapi.orth_project_host(*args, **lk)
  lib = nat.lib()
  lib.dm_debug_set_wait_guard(200_000, 1)
  lib.dm_debug_set_host_chunk(16)
  try:
    with pytest.raises(nat.NativeError, match="timed out"):
      hostapi.orth_project_host(*args, **lk)     # synchronous entry: reported by the call itself
  finally:
    lib.dm_debug_set_wait_guard(0, 0)
    lib.dm_debug_set_host_chunk(0)
  again = hostapi.orth_project_host(*args, **lk)
  assert_same(again[0], good[0], "topdown after recovery")
  assert_same(again[1], good[1], "mask after recovery")


# ---- stress: the cross-CTA protocol of the persistent kernels ---------------------------------------------------

def _gpu_digest(*tensors):
  """Order-sensitive 64-bit digest computed on the device (no D2H of the maps): sum of value bits times a
  position-dependent odd multiplier, in int64 wrap-around arithmetic."""
  acc = torch.zeros((), dtype=torch.int64, device=tensors[0].device)
  for t in tensors:
    v = t.contiguous().view(-1)
    v = v.view(torch.int32).to(torch.int64) if v.dtype == torch.float32 else v.to(torch.int64)
    idx = torch.arange(v.numel(), dtype=torch.int64, device=v.device)
    acc = acc * 1000003 + ((v + 0x9E37) * (2 * idx + 1)).sum()
  return acc


@pytest.mark.parametrize("case", ["cfg2_room_float", "cfg2_iid_float", "cfg2_room_labels", "cfg2_iid_labels",
                                  "cfg5_labels", "cfg5_float"])
def test_stress_identical_digest_over_many_launches(case):
  """>= 2000 back-to-back launches in total over the cases (config-2 room / iid, config-5 shapes; float and label
  kernels): the digest of (topdown, mask, height) never changes and no wait ever times out.  A missing
  acquire / release edge in the ticket protocol shows up here as a digest that differs once in a while."""
  cfg5 = case.startswith("cfg5")
  b, H, W, C = (8, 720, 1280, 40) if cfg5 else (64, 480, 640, 16)
  launches = 120 if cfg5 else 450
  scene = "iid" if "iid" in case else "room"
  depth, values, pose = synth.frames(scene, b, H, W, C, seed=21, device="cuda")
  labels = values.argmax(1, keepdim=True).to(torch.uint8)
  proj = dmap.MapProjector(width=W, height=H, hfov=HFOV, cam_pose=[0., 0., 0.], width_offset=200., height_offset=0.,
                           cam_pitch=PITCH, cam_height=0.88, map_res=0.03, map_width=400, map_height=400,
                           trunc_depth_min=0.15, trunc_depth_max=5.05, clip_border=10, to_global=False,
                           fill_value=dmap.NINF)
  if case.endswith("labels"):
    del values
    kw = dict(label_map=labels, num_classes=C)
  else:
    kw = dict(value_map=values)
  first = None
  digests = []
  for i in range(launches):
    top, mask, hgt = proj.orth_project(depth, cam_pose=pose, get_height_map=True, **kw)
    digests.append(_gpu_digest(top, mask, hgt[:, :1]))
    if i % 50 == 49:
      torch.cuda.synchronize()
  torch.cuda.synchronize()
  vals = torch.stack(digests).cpu().numpy()
  assert (vals == vals[0]).all(), f"{int((vals != vals[0]).sum())} of {launches} launches differ"
  nat.device_status()
