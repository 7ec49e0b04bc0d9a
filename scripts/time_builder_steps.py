#!/usr/bin/env python
"""Per-step GPU time of a bench.py builder workload (events around every step) + host enqueue time."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
key = sys.argv[1]
dev = torch.device("cuda", 0)
w = bench.BuilderWorkload(types.SimpleNamespace(scene="room"), key)
w.t = 0
w.setup(dev, 0)
w.reset_counters()
for rep in range(3):
  for _ in range(w.EPISODE if rep == 0 else 0): w.step()
  torch.cuda.synchronize()
  evs = [torch.cuda.Event(enable_timing=True) for _ in range(w.EPISODE + 1)]
  host = []
  evs[0].record()
  for i in range(w.EPISODE):
    t0 = time.perf_counter()
    w.step()
    host.append(time.perf_counter() - t0)
    evs[i + 1].record()
  torch.cuda.synchronize()
  g = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(w.EPISODE)]) * 1e3
  h = np.array(host) * 1e6
  print(key, "rep", rep, "gpu us/step: mean %.0f median %.0f p90 %.0f max %.0f | host us/step: mean %.0f median %.0f p90 %.0f max %.0f | total ms/step %.4f" % (
    g.mean(), np.median(g), np.percentile(g, 90), g.max(), h.mean(), np.median(h), np.percentile(h, 90), h.max(), evs[0].elapsed_time(evs[-1]) / w.EPISODE))
  print("   slowest steps (index: gpu us, host us):", [(int(i), int(g[i]), int(h[i])) for i in np.argsort(-g)[:6]])
