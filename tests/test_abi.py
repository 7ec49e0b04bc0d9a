"""The drop-in boundary without a GPU: include/dungeon_maps_b200.h, the ctypes binding and the
built .so agree — every declared entry point is exported, every struct has the size/offsets the
header documents — and the host-side parameter packing equals the oracle's.  No compute calls.
"""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from dungeon_maps_b200 import _native as nat, _params
from oracle import dm_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dungeon_maps_b200.h")


def _declared_functions():
  src = open(HEADER).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)     # comments
  src = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", src, flags=re.S)
  src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
  return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_binding_binds():
  assert _declared_functions() == sorted(nat.EXPORTS)


def test_library_loads_and_exports_every_declared_symbol():
  handle = ctypes.CDLL(nat.library_path()) if os.path.exists(nat.library_path()) else nat.lib()
  for name in _declared_functions():
    assert hasattr(handle, name), f"libdungeon_maps_b200.so does not export {name}"
  lib = nat.lib()
  assert lib.dm_abi_version() == nat.ABI_VERSION == 3
  assert b"sm_100a" in lib.dm_build_info()
  assert lib.dm_launch_count() >= 0


def test_library_is_standalone():
  """No torch / libcudart.so dependency: plain C ABI, cudart linked statically."""
  import subprocess
  out = subprocess.run(["ldd", nat.library_path()], capture_output=True, text=True).stdout
  assert "libtorch" not in out and "libc10" not in out and "libcudart" not in out, out


def test_struct_layouts_match_the_header():
  assert ctypes.sizeof(nat.DmStep) == 64
  assert nat.DmStep.t.offset == 36 and nat.DmStep.kind.offset == 48 and nat.DmStep.fused.offset == 52
  assert nat.PROJ_SAMPLE_WORDS * 4 == 192 == orc.PROJ_SAMPLE_DT.itemsize
  assert nat.FLOW_SAMPLE_WORDS * 4 == 192 == orc.FLOW_SAMPLE_DT.itemsize
  assert ctypes.sizeof(nat.DmProjCfg) == ctypes.sizeof(orc.ProjCfg) == 25 * 4
  assert ctypes.sizeof(nat.DmFlowCfg) == ctypes.sizeof(orc.FlowCfg) == 15 * 4
  for (n1, t1), (n2, t2) in zip(nat.DmProjCfg._fields_, orc.ProjCfg._fields_):
    assert n1 == n2 and ctypes.sizeof(t1) == ctypes.sizeof(t2)
  assert ctypes.sizeof(nat.DmFuseSource) == 3 * 8 + 2 * 8 + 4 * 4 + 4 * 8
  assert ctypes.sizeof(nat.DmFuseTarget) == 8 * 4
  # every mirrored struct against sizeof() inside the compiled library
  lib = nat.lib()
  sizes = [64, nat.PROJ_SAMPLE_WORDS * 4, ctypes.sizeof(nat.DmProjCfg), nat.FLOW_SAMPLE_WORDS * 4,
           ctypes.sizeof(nat.DmFlowCfg), ctypes.sizeof(nat.DmFuseSource), ctypes.sizeof(nat.DmFuseTarget),
           ctypes.sizeof(nat.DmBuilderCfg), ctypes.sizeof(nat.DmMapRef), ctypes.sizeof(nat.DmMergeShape)]
  assert [lib.dm_sizeof_struct(i) for i in range(len(sizes))] == sizes
  assert lib.dm_sizeof_struct(99) == -1


def test_workspace_query_needs_no_device():
  lib = nat.lib()
  cfg = nat.DmProjCfg(H=480, W=640, C=16, Mh=400, Mw=400, want_height=1)
  need = lib.dm_orth_project_workspace_bytes(ctypes.byref(cfg), 64)
  # sparse ring of up to 10 frame slots of 400*400 cells x 17 keys + slice flags + control block
  assert 4 * 160000 * 17 * 4 <= need <= 10 * 160000 * 17 * 4 + (2 << 20)
  # a single frame still gets the two slots the schedule needs
  assert lib.dm_orth_project_workspace_bytes(ctypes.byref(cfg), 1) >= 2 * 160000 * 17 * 4
  assert lib.dm_orth_project_workspace_bytes(None, 64) == 0
  assert lib.dm_orth_project_workspace_bytes(ctypes.byref(cfg), 0) == 0


def test_usage_errors_are_codes_not_crashes():
  lib = nat.lib()
  cfg = nat.DmProjCfg(H=4, W=4, C=0, Mh=4, Mw=4)
  assert lib.dm_orth_project_f32(0, 0, 0, 0, None, 1, 0, 0, 0, 0, 0, 0) == -1
  assert lib.dm_orth_project_f32(0, 0, 0, 0, ctypes.byref(cfg), 0, 0, 0, 0, 0, 0, 0) == 0   # empty batch
  assert lib.dm_orth_project_f32(0, 0, 0, 0, ctypes.byref(cfg), 1, 0, 0, 0, 0, 0, 0) == -1  # null pointers
  with pytest.raises(nat.NativeError):
    nat.check(-2, "x")


def test_no_cpu_fallback():
  if torch.cuda.is_available():
    pytest.skip("CPU-only check")
  import dungeon_maps_b200 as dmap
  with pytest.raises(nat.NativeError):
    dmap.orth_project(torch.ones(1, 1, 4, 4), None, None, [0., 0., 0.], 2., 0., 0., 0.88, 0.5, 4, 4, 2., 2., 2., 2.,
                      None, None, None, None, False)
  with pytest.raises(nat.NativeError):
    dmap.camera_affine_grid(torch.ones(1, 1, 4, 4), [0., 0., 0.], 0., 0.88, 2., 2., 2., 2.)
  with pytest.raises(nat.NativeError):
    nat.require_cuda("cpu")


def _steps_np(t: torch.Tensor):
  return np.frombuffer(t.contiguous().numpy().tobytes(), dtype=orc.STEP_DT)


@pytest.mark.parametrize("n_points", [1, 33, 44, 45, 307200])
def test_host_parameter_blocks_equal_the_oracles(n_points):
  """_params (product host code) and oracle/dm_oracle.py build the DmStep blocks independently
  (both with the reference's torch-CPU op sequence utils.py:303-327): same bytes."""
  rng = np.random.default_rng(n_points)
  b = 7
  pitch = torch.from_numpy(rng.normal(size=b).astype(np.float32) * 0.3)
  pitch[0] = 0.0005   # |angle| <= ANGLE_EPS → identity (utils.py:323-324)
  camh = torch.from_numpy(rng.uniform(0.5, 1.5, size=b).astype(np.float32))
  pose = torch.from_numpy(rng.normal(size=(b, 3)).astype(np.float32))
  pairs = [
    (_params.camera_to_local(pitch, camh, n_points), orc.to_local_steps(pitch.numpy(), camh.numpy(), n_points)),
    (_params.local_to_camera(pitch, camh, n_points), orc.to_camera_steps(pitch.numpy(), camh.numpy(), n_points)),
    (_params.local_to_global(pose, n_points), orc.to_global_steps(pose.numpy(), n_points)),
    (_params.global_to_local(pose, n_points), orc.from_global_steps(pose.numpy(), n_points)),
  ]
  for ours, theirs in pairs:
    ours = _steps_np(ours)
    for f in ("R", "t", "kind", "fused"):
      assert np.array_equal(ours[f], np.asarray(theirs)[f]), f
  assert _params.fused_for(45) and not _params.fused_for(44)   # 9*n >= 400 → MKL sgemm FMA chain


@pytest.mark.parametrize("n_points", [33, 45, 307200])
def test_c_packers_equal_the_python_blocks(n_points):
  """dm_pack_proj_samples / dm_pack_flow_samples (csrc/dm_params.cu, pure host code) against the torch / numpy
  construction of _params.py, byte for byte — including the |yaw| <= 0.001 clamp, yaw = 0, +-pi, and fast_steps."""
  rng = np.random.default_rng(7 + n_points)
  b = 19
  pose = torch.from_numpy(rng.normal(size=(b, 3)).astype(np.float32) * np.float32([2., 2., 1.5]))
  pose[0, 2] = 0.0; pose[1, 2] = 0.0009; pose[2, 2] = -0.00099; pose[3, 2] = np.pi; pose[4, 2] = -np.pi
  pose[5, 2] = 0.0010001; pose[6, 2] = np.pi / 2
  woff = torch.from_numpy(rng.normal(size=b).astype(np.float32) * 50)
  hoff = torch.from_numpy(rng.normal(size=b).astype(np.float32) * 50)
  for pitch_v in (-0.17453292, 0.0, 0.0005, 0.6):
    pitch = _params.per_sample(pitch_v, b)
    camh = _params.per_sample(0.88, b)
    assert _params._uniform(pitch) and _params._uniform(camh)
    real = lambda t: t.expand(b).clone()  # same values, not recognisably uniform: the torch / numpy path
    for to_global in (False, True):
      got, fast = _params.proj_samples(pose, pitch, camh, woff, hoff, to_global, n_points)
      want, fast_w = _params.proj_samples(pose, real(pitch), real(camh), woff, hoff, to_global, n_points)
      assert got.numpy().tobytes() == want.numpy().tobytes(), (pitch_v, to_global)
      assert fast == fast_w
    got = _params.flow_samples(pose, pitch, camh, n_points)
    want = _params.flow_samples(pose, real(pitch), real(camh), n_points)
    assert got.numpy().tobytes() == want.numpy().tobytes(), pitch_v


def test_tracked_box_guard_host_logic():
  """The bounding box fuse_topdown_maps keeps next to a map it wrote is only trusted while everything it was derived
  from is unchanged (maps._tracked_box_valid): same mask tensor at the same in-place version, same projector object
  with the same offsets / resolution / flip, global-frame map and target.  Pure host logic, no kernels."""
  import math
  import dungeon_maps_b200 as dmap
  from dungeon_maps_b200 import maps as M
  cpu = torch.device("cpu")
  proj = dmap.MapProjector(width=64, height=48, hfov=math.radians(70), cam_pose=[0., 0., 0.],
                           width_offset=torch.tensor([3.5]), height_offset=torch.tensor([2.0]), cam_pitch=0.,
                           cam_height=0.88, map_res=0.05, map_width=20, map_height=20, to_global=True)
  mask = torch.zeros((1, 1, 20, 20), dtype=torch.bool)
  hm = torch.zeros((1, 1, 20, 20))
  m = dmap.TopdownMap(topdown_map=hm, mask=mask, height_map=hm, map_projector=proj)
  assert not M._tracked_box_valid(m, proj, cpu)                       # nothing tracked yet
  m._tracked_box = M._TrackedBox(torch.zeros(5, dtype=torch.int64), mask, proj)
  assert M._tracked_box_valid(m, proj.clone(cam_pose=[1., 0., 0.5]), cpu)   # the target's pose does not matter
  assert not M._tracked_box_valid(m, proj.clone(to_global=False), cpu)     # local target: points get rotated
  assert not M._tracked_box_valid(m, proj.clone(map_res=0.1), cpu)         # other bin size
  assert not M._tracked_box_valid(m, proj, torch.device("cuda", 0))        # box lives on another device
  other = dmap.TopdownMap(topdown_map=hm, mask=mask.clone(), height_map=hm, map_projector=proj)
  other._tracked_box = m._tracked_box
  assert not M._tracked_box_valid(other, proj, cpu)                        # not the mask the box was derived from
  proj.width_offset = torch.tensor([4.5])                                  # projector edited in place
  assert not M._tracked_box_valid(m, proj, cpu)
  proj.width_offset = torch.tensor([3.5])
  assert M._tracked_box_valid(m, proj, cpu)
  mask[0, 0, 0, 0] = True                                                  # mask edited in place
  assert not M._tracked_box_valid(m, proj, cpu)


def test_canvas_size_classes():
  """World canvases are allocated in size classes (maps._canvas_cap) so that the caching allocator can reuse the block
  of the slightly smaller previous map, and the MapBuilder step takes canvases of the OLD map's class before the
  bounding box is known: a class must hold the map, waste at most a quarter, and never shrink as the map grows."""
  from dungeon_maps_b200 import maps as dmaps
  prev = 0
  for n in list(range(1 << 18, (1 << 18) + 4096, 37)) + [10 ** 6, 32 * 2204 * 2204, 32 * 2205 * 2204, (1 << 28) - 1,
                                                         1 << 28, (1 << 28) + 1]:
    cap = dmaps._canvas_cap(n)
    assert n <= cap <= n + n // 4 + 8, (n, cap)
    top = 1 << (n - 1).bit_length()
    assert cap in (top // 2 + k * (top // 8) for k in range(1, 5)), (n, cap)
  for n in range(1 << 18, 1 << 20, 4099):
    cap = dmaps._canvas_cap(n)
    assert cap >= prev
    prev = cap


def test_workspace_of_the_height_map_path():
  """Depth-only calls (C = 0) get a key plane per frame, up to 64 per launch (csrc/dm_project.cu: hmap_* kernels); the
  workspace query says so without a device."""
  lib = nat.lib()
  cfg = nat.DmProjCfg(H=480, W=640, C=0, Mh=400, Mw=400)
  plane = 400 * 400 * 4
  one, many, more = (lib.dm_orth_project_workspace_bytes(ctypes.byref(cfg), b) for b in (1, 64, 200))
  assert many - one >= 62 * plane, "a plane per frame"
  assert 0 <= more - many < 4096, "at most 64 planes (longer batches are chunked); only the control block grows"
  cfg.C = 16
  assert lib.dm_orth_project_workspace_bytes(ctypes.byref(cfg), 64) < 11 * 400 * 400 * 17 * 4 + (8 << 20), \
      "float planes keep the 10-slot ring"
