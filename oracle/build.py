"""ORACLE — test infrastructure only.  Compiles oracle/dm_oracle.c into
oracle/_build/libdm_oracle.so with gcc (no CUDA, no torch).

-ffp-contract=off: the compiler must not fuse a*b+c on its own; the C code
calls fmaf() exactly where the reference's sgemm fused (see dm_oracle.c).
-mfma: fmaf() becomes one vfmadd instruction instead of a libm call.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dm_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdm_oracle.so")


def build(force: bool = False) -> str:
  os.makedirs(OUT_DIR, exist_ok=True)
  hdr = os.path.join(HERE, "..", "include", "dungeon_maps_b200.h")
  if not force and os.path.exists(LIB):
    newest = max(os.path.getmtime(SRC), os.path.getmtime(hdr))
    if os.path.getmtime(LIB) >= newest:
      return LIB
  cmd = [
    "gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-shared", "-fPIC",
    "-Wall", "-o", LIB, SRC, "-lm",
  ]
  subprocess.run(cmd, check=True)
  return LIB


if __name__ == "__main__":
  print(build(force=True))
