"""Loader for the upstream pure-Python reference (TEST INFRASTRUCTURE ONLY).

Imports `starks` from the unmodified copy staged under baseline/_ref/ (git-ignored; made by
baseline/stage_ref.py, shipped to the GPU box by gpurun) or from /root/reference (authoring
container only) and applies the two-line restoration SURVEY.md App. B documents:
the multiplicative-subgroup FRI in starks/fri.py:176-366 is commented out at HEAD, so
the leading '#' of those lines is dropped at load time and `FRI = SmoothSubgroupFRI`
is appended.  No reference source is copied into this repository; the module source
is read, patched in memory and exec'd.

Only oracle/gen_golden.py, tests (skipped when no reference tree is present) and bench.py's
CPU-baseline leg may import this file.
"""
import importlib.util
import io
import os
import sys
import contextlib

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _locate():
  """STARKS_REFERENCE, else the copy staged by baseline/stage_ref.py (the only one that exists on
  the GPU box), else the authoring container's /root/reference."""
  env = os.environ.get("STARKS_REFERENCE")
  if env:
    return env
  staged = os.path.join(_REPO, "baseline", "_ref")
  if os.path.isdir(os.path.join(staged, "starks")):
    return staged
  return "/root/reference"


REF_ROOT = _locate()


def available() -> bool:
  return os.path.isdir(os.path.join(REF_ROOT, "starks"))


def load():
  """Returns the reference `starks` package with starks.fri restored."""
  if not available():
    raise RuntimeError("reference tree not present at %s" % REF_ROOT)
  if REF_ROOT not in sys.path:
    sys.path.insert(0, REF_ROOT)
  if "starks.fri" in sys.modules and hasattr(sys.modules["starks.fri"], "SmoothSubgroupFRI"):
    import starks
    return starks
  import warnings
  warnings.filterwarnings("ignore", category=SyntaxWarning)
  import starks  # noqa: F401  (package __init__ is empty)
  fri_path = os.path.join(REF_ROOT, "starks", "fri.py")
  with open(fri_path) as fh:
    lines = fh.read().split("\n")
  out = []
  for i, line in enumerate(lines, start=1):
    if i >= 176 and line.startswith("#"):
      line = line[1:]
    out.append(line)
  src = "\n".join(out) + "\nFRI = SmoothSubgroupFRI\n"
  spec = importlib.util.spec_from_loader("starks.fri", loader=None, origin=fri_path)
  mod = importlib.util.module_from_spec(spec)
  mod.__file__ = fri_path
  sys.modules["starks.fri"] = mod
  exec(compile(src, fri_path, "exec"), mod.__dict__)
  starks.fri = mod
  return starks


@contextlib.contextmanager
def quiet():
  """The reference prints progress (and whole coefficient lists); swallow it."""
  buf = io.StringIO()
  with contextlib.redirect_stdout(buf):
    yield buf
