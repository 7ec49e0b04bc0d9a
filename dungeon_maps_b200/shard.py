"""Batch sharding across the GPUs of one box.

Every frame of orth_project / camera_affine_grid is independent and every environment of a
MapBuilder owns its world map, so the batch dimension is cut into contiguous ranges, one per
rank (one process per GPU), and NO collective sits on the data path.  torch.distributed is used
only for the plumbing around it: barriers, max-over-ranks timing, and an optional gather of the
finished maps onto one rank for callers that want them in one place.
"""
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
  """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
  return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
          int(os.environ.get("LOCAL_RANK", "0")))


def frame_range(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
  """Contiguous [begin, end) of the frames rank owns; sizes differ by at most one, earlier
  ranks take the larger shares; ranks beyond the frame count get an empty range."""
  if n_frames < 0 or world_size < 1 or not 0 <= rank < world_size:
    raise ValueError(f"bad shard request: n_frames={n_frames} rank={rank} world_size={world_size}")
  base, extra = divmod(n_frames, world_size)
  begin = rank * base + min(rank, extra)
  return begin, begin + base + (1 if rank < extra else 0)


def frame_ranges(n_frames: int, world_size: int) -> List[Tuple[int, int]]:
  return [frame_range(n_frames, r, world_size) for r in range(world_size)]


def take(rank: int, world_size: int, *tensors):
  """The rank's slice along dim 0 of every tensor (None passes through)."""
  out = []
  for t in tensors:
    if t is None:
      out.append(None)
      continue
    lo, hi = frame_range(t.shape[0], rank, world_size)
    out.append(t[lo:hi])
  return out[0] if len(out) == 1 else tuple(out)


def _active() -> bool:
  return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def barrier(device: Optional[torch.device] = None) -> None:
  if _active():
    dist.barrier()
  if device is not None and torch.device(device).type == "cuda":
    torch.cuda.synchronize(device)


def _reduce(x: float, op, device) -> float:
  if not _active():
    return float(x)
  t = torch.tensor([x], dtype=torch.float64, device=device if device is not None else "cpu")
  dist.all_reduce(t, op=op)
  return float(t.item())


def max_over_ranks(x: float, device=None) -> float:
  """Multi-GPU timings are the max over ranks of per-rank device times."""
  return _reduce(x, dist.ReduceOp.MAX, device)


def sum_over_ranks(x: float, device=None) -> float:
  return _reduce(x, dist.ReduceOp.SUM, device)


def gather_frames(local: torch.Tensor, n_frames: int, dst: int = 0) -> Optional[torch.Tensor]:
  """Off the data path: reassembles the (n_frames, ...) result on rank dst from every rank's
  frame_range slice.  Returns None on the other ranks."""
  if not _active():
    return local
  rank, ws = dist.get_rank(), dist.get_world_size()
  ranges = frame_ranges(n_frames, ws)
  lo, hi = ranges[rank]
  assert local.shape[0] == hi - lo, f"rank {rank} holds {local.shape[0]} frames, owns {hi - lo}"
  biggest = max(h - l for l, h in ranges)
  padded = local.new_zeros((biggest,) + tuple(local.shape[1:]))
  padded[:hi - lo] = local
  bufs = [torch.empty_like(padded) for _ in range(ws)] if rank == dst else None
  dist.gather(padded, bufs, dst=dst)
  if rank != dst:
    return None
  return torch.cat([b[:h - l] for b, (l, h) in zip(bufs, ranges)], dim=0)


def bind_to_gpu_numa_node(local_rank: int) -> Optional[int]:
  """Host-buffer entries are PCIe-bound: pin this process (and therefore the pinned staging memory it touches
  first) to the CPUs of the NUMA node the rank's GPU hangs off, so that N ranks do not all stream through one
  socket's memory.  Returns the node, or None when the topology cannot be read (then nothing is changed)."""
  try:
    props = torch.cuda.get_device_properties(local_rank)
    bdf = f"{int(props.pci_domain_id):04x}:{int(props.pci_bus_id):02x}:{int(props.pci_device_id):02x}.0"
    with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
      node = int(f.read().strip())
    if node < 0:
      return None
    with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
      cpus = set()
      for part in f.read().strip().split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    allowed = cpus & os.sched_getaffinity(0)
    if not allowed:
      return None
    os.sched_setaffinity(0, allowed)
    return node
  except Exception:
    return None

