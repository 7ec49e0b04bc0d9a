#!/usr/bin/env python
"""Replays bench.run_extras' builder -> builder_fixed sequence with per-step events."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
dev = torch.device("cuda", 0)
args = types.SimpleNamespace(scene="room", workload="proj")
shared = None
for key in (["proj_labels", "flow"] if "--full" in sys.argv else []) + ["builder", "builder_fixed"]:
  wl = bench.make_workload(args, key)
  if isinstance(wl, bench.BuilderWorkload):
    wl.setup(dev, 0, frames=shared); shared = wl.frames
  else:
    wl.setup(dev, 0)
  for _ in range(wl.EPISODE if isinstance(wl, bench.BuilderWorkload) else 3): wl.step()
  if hasattr(wl, "reset_counters"):
    wl.t = 0; wl.reset_counters()
  torch.cuda.synchronize()
  n = 100
  evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
  host = []
  evs[0].record()
  for i in range(n):
    t0 = time.perf_counter(); wl.step(); host.append(time.perf_counter() - t0); evs[i + 1].record()
  torch.cuda.synchronize()
  g = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(n)]) * 1e3
  h = np.array(host) * 1e6
  print(key, "total ms/step %.4f gpu median %.0f max %.0f host median %.0f max %.0f" % (evs[0].elapsed_time(evs[-1]) / n, np.median(g), g.max(), np.median(h), h.max()),
        "slowest:", [(int(i), int(g[i]), int(h[i])) for i in np.argsort(-g)[:4]])
  if key == "builder_fixed": shared = None
  wl = None
  torch.cuda.empty_cache()
