#!/usr/bin/env python
"""Experiment helper: cProfile of the host side of MapBuilder.step on the config-4 workload (needs a GPU)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

class A: pass
args = A(); args.scene = "room"
wl = bench.BuilderWorkload(args, sys.argv[1] if len(sys.argv) > 1 else "builder")
wl.setup(torch.device("cuda", 0), 0)
for _ in range(20):
  wl.step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
  wl.step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
