#!/usr/bin/env python
"""Prints the headline numbers of a bench.py JSON line: usage scripts/show_bench.py <file.json>"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.0f %s  ms/step %.4f  frac %.3f  e2e %.0f  cpu %s  launches %s" % (
  d["value"], d["unit"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"],
  (d.get("cpu_baseline") or {}).get("value"), d.get("gpu_launches")))
for k, v in (d.get("extra") or {}).items():
  print("  extra %-14s value %-12s ms/step %-8s frac %s" % (k, round(v["value"]) if "value" in v else v,
        round(v["ms_per_step"], 4) if v.get("ms_per_step") else None, round(v["roofline_frac"], 3) if v.get("roofline_frac") else None))
