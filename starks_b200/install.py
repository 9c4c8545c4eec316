"""install(): rebinds the upstream `starks` package's hot-path functions onto this library,
so that an unmodified starks/stark.py and the upstream tests run on the GPU.

    import starks_b200.install as shim; shim.install()

Rebound (module attribute and every `from ... import` alias already bound in starks.stark,
starks.fri, starks.utils -- stark.py:4-17, fri.py:6-8,17):
  starks.fft.{fft_1d, mul_polys, NonBinaryFFT}
  starks.merkle_tree.{merkelize, merkelize_polynomial_evaluations}
  starks.utils.get_power_cycle
  starks.fri.{SmoothSubgroupFRI, FRI}      (restores the class upstream HEAD comments out)
  starks.stark.STARK.mk_proof               (device-resident prover; same proof object)
Everything else (mk_branch, verify_branch, verify_proof, the AIR classes ...) is left as is:
those are list indexing / O(log n) host work.  There is no CPU fallback: after install()
the rebound functions require the CUDA library."""
import importlib
import sys


def install():
  from . import fft as bfft, merkle_tree as bmt, fri as bfri, utils as butils, stark as bstark
  try:
    fft = importlib.import_module("starks.fft")
    mt = importlib.import_module("starks.merkle_tree")
    utils = importlib.import_module("starks.utils")
  except ImportError as e:  # pragma: no cover
    raise ImportError("install() needs the upstream `starks` package on sys.path: %s" % e)

  def _polys_nbfft(field, root_of_unity):
    # keep returning the upstream Polynomial type from inv_fft
    from starks.polynomial import polynomials_over
    obj = bfft.NonBinaryFFT(field, root_of_unity)
    obj.polysOver = polynomials_over(field).factory
    return obj

  rebinds = {
      "starks.fft": {"fft_1d": bfft.fft_1d, "mul_polys": bfft.mul_polys, "NonBinaryFFT": _polys_nbfft},
      "starks.merkle_tree": {"merkelize": bmt.merkelize,
                             "merkelize_polynomial_evaluations": bmt.merkelize_polynomial_evaluations},
      "starks.utils": {"get_power_cycle": butils.get_power_cycle},
  }
  for modname, table in rebinds.items():
    mod = sys.modules[modname]
    for name, fn in table.items():
      setattr(mod, name, fn)
  # the multiplicative FRI does not exist at upstream HEAD: provide it
  try:
    fri = importlib.import_module("starks.fri")
  except ImportError:
    fri = None
  if fri is not None:
    fri.SmoothSubgroupFRI = bfri.SmoothSubgroupFRI
    fri.FRI = bfri.FRI
  # aliases bound by `from x import y` in already-imported modules
  for modname in ("starks.stark", "starks.fri", "starks.utils", "starks.compression"):
    mod = sys.modules.get(modname)
    if mod is None:
      continue
    for table in rebinds.values():
      for name, fn in table.items():
        if hasattr(mod, name):
          setattr(mod, name, fn)
    if hasattr(mod, "FRI"):
      mod.FRI = bfri.FRI
  stark = sys.modules.get("starks.stark")
  if stark is not None:
    def mk_proof(self, witness, boundary):
      dev = bstark.STARK(self.field, self.steps, self.extension_factor, self.width, self.step_polys)
      return dev.mk_proof(witness, boundary)
    stark.STARK.mk_proof = mk_proof
  return True
