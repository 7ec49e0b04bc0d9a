"""The N > 1 path on CPU: two processes over gloo run the sharding plumbing bench.py and a
multi-GPU caller use (dungeon_maps_b200.shard) — contiguous frame ranges, no data-path collective,
max/sum-over-ranks, and the off-path gather.  The per-shard compute stands in for the CUDA call
with the oracle (tests may use it as the checker's engine): sharded == unsharded, bit for bit.
"""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dungeon_maps_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_ranges_partition_the_batch():
  for n in (0, 1, 2, 7, 64, 4096, 4099):
    for ws in (1, 2, 3, 4, 8):
      r = shard.frame_ranges(n, ws)
      assert r[0][0] == 0 and r[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(r, r[1:]))           # contiguous, no gaps, no overlap
      sizes = [h - l for l, h in r]
      assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
  assert shard.frame_range(64, 3, 8) == (24, 32)
  with pytest.raises(ValueError):
    shard.frame_range(8, 2, 2)


def test_single_process_helpers_are_identity():
  assert shard.max_over_ranks(3.5) == 3.5 and shard.sum_over_ranks(2.0) == 2.0
  t = torch.arange(12).view(6, 2)
  assert shard.gather_frames(t, 6) is t
  a, b, c = shard.take(1, 4, t, None, t[:5])
  assert a.tolist() == t[2:4].tolist() and b is None and c.tolist() == t[2:3].tolist()


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world_size, port, n_frames, q):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                    LOCAL_RANK=str(rank))
  dist.init_process_group("gloo", rank=rank, world_size=world_size)
  try:
    from oracle import dm_oracle as orc
    assert shard.world() == (rank, world_size, rank)
    H, W, C = 24, 32, 3
    intr = orc.intrinsics(W, H, math.radians(70))
    kw = dict(map_res=0.1, map_width=40, map_height=40, focal_x=intr["fx"], focal_y=intr["fy"], center_x=intr["cx"],
              center_y=intr["cy"], trunc_depth_min=0.15, trunc_depth_max=5.05, trunc_height_max=None, clip_border=1,
              to_global=True, flip_h=True, fill_value=-np.inf, reduction=None, get_height_map=True)
    depth, values, pose = synth.frames("iid", n_frames, H, W, C, seed=11)       # same on every rank
    d, v, p = shard.take(rank, world_size, depth, values, pose)                  # this rank's environments
    lo, hi = shard.frame_range(n_frames, rank, world_size)
    assert d.shape[0] == hi - lo
    top, mask, hgt = orc.orth_project(d.numpy(), v.numpy(), None, p.numpy(), 20., 20., math.radians(-10), 0.88, **kw)
    shard.barrier()
    t_max = shard.max_over_ranks(float(rank + 1))          # the bench's timing reduction
    n_sum = shard.sum_over_ranks(float(hi - lo))
    g_top = shard.gather_frames(torch.from_numpy(top), n_frames)
    g_mask = shard.gather_frames(torch.from_numpy(mask.astype(np.uint8)), n_frames)
    g_hgt = shard.gather_frames(torch.from_numpy(hgt), n_frames)
    if rank == 0:
      want = orc.orth_project(depth.numpy(), values.numpy(), None, pose.numpy(), 20., 20., math.radians(-10), 0.88, **kw)
      ok = (np.array_equal(g_top.numpy(), want[0]) and np.array_equal(g_mask.numpy(), want[1].astype(np.uint8))
            and np.array_equal(g_hgt.numpy(), want[2]))
      q.put(("ok" if ok else "mismatch", t_max, n_sum))
    else:
      assert g_top is None and g_mask is None
      q.put(("ok", t_max, n_sum))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_two_ranks_sharded_equals_unsharded(n_frames):
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for status, t_max, n_sum in results:
    assert status == "ok"
    assert t_max == 2.0 and n_sum == float(n_frames)
