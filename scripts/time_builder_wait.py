#!/usr/bin/env python
"""How the host time of a MapBuilder.step (bench config 4) splits: before the plot call, the plot call (queues work),
bookkeeping behind it, the wait for the bounding box, the merge call, the rest."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from dungeon_maps_b200 import _native as nat
lib = nat.lib()
T = {"plot": 0.0, "wait": 0.0, "merge": 0.0}
class Wrap:
  def __init__(self, real): self.real = real
  def __getattr__(self, name):
    f = getattr(self.real, name)
    key = {"dm_builder_plot_prefill": "plot", "dm_builder_plot_wait": "wait", "dm_builder_merge": "merge"}.get(name)
    if key is None: return f
    def g(*a):
      t0 = time.perf_counter(); r = f(*a); T[key] += time.perf_counter() - t0; return r
    return g
w = bench.BuilderWorkload(types.SimpleNamespace(scene="room"), "builder")
w.t = 0; w.setup(torch.device("cuda", 0), 0); w.reset_counters()
for _ in range(w.EPISODE): w.step()
torch.cuda.synchronize()
real_lib = nat.lib
nat.lib = lambda: Wrap(real_lib())
import dungeon_maps_b200.maps as M
t0 = time.perf_counter()
for _ in range(w.EPISODE): w.step()
torch.cuda.synchronize()
tot = time.perf_counter() - t0
n = w.EPISODE
print("per step us: total %.0f | plot call %.0f | wait for the box %.0f | merge call %.0f | other python %.0f" % (
  tot / n * 1e6, T["plot"] / n * 1e6, T["wait"] / n * 1e6, T["merge"] / n * 1e6, (tot - sum(T.values())) / n * 1e6))
