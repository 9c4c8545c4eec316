#!/usr/bin/env python
"""Stages the UNMODIFIED upstream `starks` package under baseline/_ref/ (git-ignored, but
shipped to the GPU box by gpurun like the built .so files), so that on a box without
/root/reference

  * the drop-in tests can run the upstream modules and the upstream unit tests through
    starks_b200.install (tests/test_gpu_upstream.py),
  * bench.py can time the pure-Python reference beside the GPU numbers (cpu_baseline,
    kind "reference").

Nothing is patched on disk: the one restoration the modp path needs (SURVEY.md App. B:
starks/fri.py:176-366 is commented out at HEAD) is applied in memory at load time by
oracle/pyref.py.  No reference source enters the git history.

    python baseline/stage_ref.py [--src /root/reference] [--force]
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def stage(src="/root/reference", force=False):
  """Copies <src>/starks (sources and its test directory) to baseline/_ref/starks.  Returns the
  staged root, or None when the source tree is absent (GPU box: the prebuilt copy is used)."""
  pkg = os.path.join(src, "starks")
  if not os.path.isdir(pkg):
    return DST if os.path.isdir(os.path.join(DST, "starks")) else None
  out = os.path.join(DST, "starks")
  if os.path.isdir(out) and not force:
    # refresh only when the source is newer than the staged copy
    newest = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(pkg) for f in fs)
    if os.path.getmtime(out) >= newest:
      return DST
  if os.path.isdir(out):
    shutil.rmtree(out)
  os.makedirs(DST, exist_ok=True)
  shutil.copytree(pkg, out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
  os.utime(out, None)
  return DST


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--src", default="/root/reference")
  ap.add_argument("--force", action="store_true")
  a = ap.parse_args()
  r = stage(a.src, a.force)
  print("staged:", r)
  sys.exit(0 if r else 1)
