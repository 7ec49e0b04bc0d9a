"""ORACLE — test infrastructure only.  Imports the UNMODIFIED reference: from /root/reference in the authoring
container (fixture generation, oracle/make_golden.py), or from baseline/_ref — the `pip install --target` copy of the
same tree that travels to the GPU box — for the `--impl reference` arm of bench.py, the only thing on that box that
may import this module.

The reference imports `torch_scatter` at module top (utils.py:16), which is not
installed.  A stand-in is registered first: scatter_{max,min,add,mul} with an
`out=` tensor are `out.scatter_reduce_(dim, index.expand_as(src), src, reduce,
include_self=True)`, which for max/min is semantically identical to
torch_scatter (the existing `out` participates, the result is order-independent).
scatter_mean is restated from torch_scatter's published source (unpinned, see _mean).
"""
import sys
import types

import torch

import os

REFERENCE_ROOT = "/root/reference"
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_root():
  """Where an importable copy of the reference lives, or None."""
  for root in (REFERENCE_ROOT, INSTALLED_ROOT):
    if os.path.isdir(os.path.join(root, "dungeon_maps")):
      return root
  return None


def _make(reduce: str):
  def scatter(src, index, dim=-1, out=None, dim_size=None):
    assert out is not None, "the reference always passes out="
    out.scatter_reduce_(dim, index.expand_as(src), src, reduce=reduce, include_self=True)
    return out, None
  return scatter


def _mean(src, index, dim=-1, out=None, dim_size=None):
  """torch_scatter.scatter_mean restated from its published source (torch_scatter/scatter.py, 2.x: scatter_sum
  into out, scatter_sum of ones into a fresh count, count.clamp_(1), out.true_divide_(count)).  The package is not
  installed here, so this stand-in is NOT pinned against the real one: parity of Reduction.mean is unpinned."""
  assert out is not None, "the reference always passes out="
  idx = index.expand_as(src)
  out.scatter_add_(dim, idx, src)
  count = torch.zeros_like(out).scatter_add_(dim, idx, torch.ones_like(src))
  count.clamp_(min=1)
  out.true_divide_(count)
  return out


def load_reference():
  """Returns the reference package `dungeon_maps` (v0.0.3a1)."""
  if "torch_scatter" not in sys.modules:
    ts = types.ModuleType("torch_scatter")
    ts.scatter_max = _make("amax")
    ts.scatter_min = _make("amin")
    ts.scatter_add = _make("sum")
    ts.scatter_mul = _make("prod")
    ts.scatter_mean = _mean
    sys.modules["torch_scatter"] = ts
  root = reference_root()
  if root is None:
    raise ImportError("the reference is neither at /root/reference nor installed under baseline/_ref")
  if root not in sys.path:
    sys.path.insert(0, root)
  import dungeon_maps  # noqa: E402
  return dungeon_maps


def per_sample(fn, batch: int, batched_args: dict, shared_args: dict):
  """The reference raises for batch > 1 (utils.py:311-316 stacks a (batch,) zeros
  tensor with (1,) axis components).  Batched semantics are therefore defined as
  the reference applied to each sample alone, outputs concatenated."""
  outs = []
  for i in range(batch):
    args = {k: (None if v is None else v[i:i + 1]) for k, v in batched_args.items()}
    out = fn(**args, **shared_args)
    outs.append(out if isinstance(out, tuple) else (out,))
  cat = tuple(torch.cat([o[j] for o in outs], dim=0) for j in range(len(outs[0])))
  return cat if len(cat) > 1 else cat[0]
