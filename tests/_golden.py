"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the unmodified reference)."""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
  def __init__(self, name):
    self.name = name
    self._z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    self.meta = json.loads(bytes(self._z["meta"]).decode())

  def __contains__(self, k):
    return k in self._z.files

  def __getitem__(self, k):
    return self._z[k]

  def get(self, k, default=None):
    return self._z[k] if k in self._z.files else default

  @property
  def kwargs(self):
    return dict(self.meta.get("kwargs", {}))


def names(prefix):
  return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith(prefix) and f.endswith(".npz"))


def sha(a) -> str:
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_same(got, want, what=""):
  """Bit-exact for ints/bools; for floats: identical values with NaN == NaN and -0 == +0
  (the reference's own results carry no information in the sign of zero)."""
  got = np.asarray(got)
  want = np.asarray(want)
  assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
  if got.dtype.kind == "f" or want.dtype.kind == "f":
    ok = (got == want) | (np.isnan(got) & np.isnan(want))
  else:
    ok = got == want
  if not ok.all():
    bad = np.argwhere(~ok)
    i = tuple(bad[0])
    raise AssertionError(f"{what}: {len(bad)} / {ok.size} elements differ; first at {i}: got {got[i]!r} want {want[i]!r}")
