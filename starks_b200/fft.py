"""Drop-in for the multiplicative-group part of starks/fft.py: NonBinaryFFT (:256-272),
fft_1d (:316-331) and mul_polys (:334-345), computed by libstarks_b200 (stk_ntt,
stk_mul_polys).  Same call signatures, same value conventions: inputs are ints or field
elements, outputs are lists of instances of `field`; inv_fft returns a Polynomial with
trailing zeros stripped.

Deviation (documented in DESIGN.md): an input longer than the order of the root raises
IndexError; the reference only raises it beyond twice the order and returns a partially
filled list in between (starks/fft.py:303-314 indexes whatever it was given)."""
from typing import List

import numpy as np

from .engine import default_engine
from .limbs import ints_to_limbs, limbs_to_ints
from .modp import element_to_int
from .polynomial import polynomials_over
from .utils import multiplicative_order


def _to_limbs(vals, p):
  if isinstance(vals, np.ndarray) and vals.dtype == np.uint32 and vals.ndim == 2 and vals.shape[1] == 8:
    return vals
  ints = [element_to_int(v) % p for v in vals]
  return ints_to_limbs(ints) if ints else np.zeros((0, 8), np.uint32)


def fft_1d(field, vals, modulus, root_of_unity, inv=False, engine=None):
  """Computes FFT for one dimensional inputs (starks/fft.py:316-331)."""
  p = int(modulus)
  root = element_to_int(root_of_unity) % p
  n = multiplicative_order(root, p)
  if len(vals) > n:
    raise IndexError("list index out of range")
  eng = engine or default_engine()
  eng.set_field(p)
  out = eng.ntt_host(_to_limbs(vals, p).reshape(1, -1, 8), n, root, inverse=inv)[0]
  return [field(v) for v in limbs_to_ints(out)]


def mul_polys(a, b, root_of_unity, engine=None):
  """Multiply polynomials by converting to fourier space (starks/fft.py:334-345; the
  inverse transform is NOT scaled by 1/N, exactly as upstream)."""
  field = type(root_of_unity)
  p = field.p
  root = element_to_int(root_of_unity) % p
  n = multiplicative_order(root, p)
  if len(a) > n or len(b) > n:
    raise IndexError("list index out of range")
  eng = engine or default_engine()
  eng.set_field(p)
  out = eng.mul_polys(_to_limbs(a, p), _to_limbs(b, p), n, root)
  return [field(v) for v in limbs_to_ints(out)]


class FFT(object):
  """Abstract class that specifies a FFT solver (starks/fft.py:9-19)."""

  def fft(self, poly):
    raise NotImplementedError

  def inv_fft(self, values):
    raise NotImplementedError


class NonBinaryFFT(FFT):
  """FFT that works for finite fields which don't have characteristic 2
  (starks/fft.py:256-272)."""

  def __init__(self, field, root_of_unity, engine=None):
    self.field = field
    self.root_of_unity = root_of_unity
    self.polysOver = polynomials_over(self.field).factory
    self._engine = engine

  def fft(self, poly) -> List:
    coeffs = poly.coefficients if hasattr(poly, "coefficients") else list(poly)
    return fft_1d(self.field, coeffs, self.field.p, self.root_of_unity, inv=False, engine=self._engine)

  def inv_fft(self, values):
    coeffs = fft_1d(self.field, values, self.field.p, self.root_of_unity, inv=True, engine=self._engine)
    return self.polysOver(coeffs)
