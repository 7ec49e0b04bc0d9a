#!/usr/bin/env python
"""Where the host time of a camera_affine_grid call goes (bench.py's flow workload: fresh pose deltas every step)."""
import cProfile, os, pstats, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
w = bench.FlowWorkload(types.SimpleNamespace(scene="room"))
w.setup(torch.device("cuda", 0), 0)
for _ in range(10): w.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): w.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("enqueue %.1f us/step, with drain %.1f us/step" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): w.step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
