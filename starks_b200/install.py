"""install(): rebinds the upstream `starks` package's hot-path functions onto this library,
so that an unmodified starks/stark.py and the upstream tests run on the GPU.

    import starks_b200.install as shim
    shim.install()          # ... upstream code now reaches CUDA ...
    shim.uninstall()        # restores every attribute install() touched

Rebound (module attribute and every `from ... import` alias already bound in starks.stark,
starks.fri, starks.utils, starks.compression -- stark.py:4-17, fri.py:6-8,17):
  starks.fft.{fft_1d, mul_polys, NonBinaryFFT}
  starks.merkle_tree.{merkelize, merkelize_polynomial_evaluations}
  starks.utils.get_power_cycle
  starks.fri.{SmoothSubgroupFRI, FRI}      (the class upstream HEAD comments out)
  starks.stark.STARK.mk_proof               (replace_prover=True only: device-resident prover,
                                             same proof object)
With replace_prover=False the UPSTREAM STARK.mk_proof body (stark.py:233-279) runs as written
-- coefficient-form constructions in Python -- and only its transforms, commitments and FRI
go to the GPU: the function-level drop-in of SURVEY.md 8(b).
Everything else (mk_branch, verify_branch, verify_proof, the AIR classes ...) is left as is:
those are list indexing / O(log n) host work.  There is no CPU fallback: after install()
the rebound functions require the CUDA library."""
import importlib
import sys

_saved = []          # (object, attribute name, previous value or _MISSING), in rebinding order
_MISSING = object()


def _rebind(obj, name, value):
  _saved.append((obj, name, getattr(obj, name, _MISSING)))
  setattr(obj, name, value)


def installed():
  return bool(_saved)


def uninstall():
  """Puts back every attribute install() replaced (newest first)."""
  while _saved:
    obj, name, prev = _saved.pop()
    if prev is _MISSING:
      try:
        delattr(obj, name)
      except AttributeError:
        pass
    else:
      setattr(obj, name, prev)


def install(replace_prover=True, engine=None):
  """engine: the Engine the rebound functions use (default: the process-wide one)."""
  from . import fft as bfft, merkle_tree as bmt, fri as bfri, utils as butils, stark as bstark
  if _saved:
    uninstall()
  try:
    importlib.import_module("starks.fft")
    importlib.import_module("starks.merkle_tree")
    importlib.import_module("starks.utils")
  except ImportError as e:  # pragma: no cover
    raise ImportError("install() needs the upstream `starks` package on sys.path: %s" % e)

  def _polys_nbfft(field, root_of_unity):
    # keep returning the upstream Polynomial type from inv_fft
    from starks.polynomial import polynomials_over
    obj = bfft.NonBinaryFFT(field, root_of_unity, engine=engine) if engine is not None else bfft.NonBinaryFFT(
        field, root_of_unity)
    obj.polysOver = polynomials_over(field).factory
    return obj

  def _with_engine(fn):
    if engine is None:
      return fn

    def bound(*a, **kw):
      kw.setdefault("engine", engine)
      return fn(*a, **kw)
    bound.__name__, bound.__doc__ = fn.__name__, fn.__doc__
    return bound

  _FRI = bfri.SmoothSubgroupFRI
  if engine is not None:
    class _FRI(bfri.SmoothSubgroupFRI):  # noqa: F811
      def __init__(self, field, engine_=None):
        super().__init__(field, engine=engine_ if engine_ is not None else engine)

  rebinds = {
      "starks.fft": {"fft_1d": _with_engine(bfft.fft_1d), "mul_polys": _with_engine(bfft.mul_polys),
                     "NonBinaryFFT": _polys_nbfft},
      "starks.merkle_tree": {"merkelize": _with_engine(bmt.merkelize),
                             "merkelize_polynomial_evaluations": _with_engine(bmt.merkelize_polynomial_evaluations)},
      "starks.utils": {"get_power_cycle": butils.get_power_cycle},
  }
  for modname, table in rebinds.items():
    mod = sys.modules[modname]
    for name, fn in table.items():
      _rebind(mod, name, fn)
  # the multiplicative FRI does not exist at upstream HEAD: provide it
  try:
    fri = importlib.import_module("starks.fri")
  except ImportError:
    fri = None
  if fri is not None:
    _rebind(fri, "SmoothSubgroupFRI", _FRI)
    _rebind(fri, "FRI", _FRI)
  # aliases bound by `from x import y` in already-imported modules
  for modname in ("starks.stark", "starks.fri", "starks.utils", "starks.compression"):
    mod = sys.modules.get(modname)
    if mod is None:
      continue
    for table in rebinds.values():
      for name, fn in table.items():
        if hasattr(mod, name):
          _rebind(mod, name, fn)
    if hasattr(mod, "FRI") and mod is not fri:
      _rebind(mod, "FRI", _FRI)
  stark = sys.modules.get("starks.stark")
  if stark is not None and replace_prover:
    def mk_proof(self, witness, boundary):
      dev = bstark.STARK(self.field, self.steps, self.extension_factor, self.width, self.step_polys, engine=engine)
      return dev.mk_proof(witness, boundary)
    _rebind(stark.STARK, "mk_proof", mk_proof)
  return True
