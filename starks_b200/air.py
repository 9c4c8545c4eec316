"""Trace generation of starks/air.py: get_computational_trace (:31-52) and the witness
transposition of AIR.generate_witness (:124).

The recurrence state[i+1] = step_polys(state[i]) is one sequential 256-bit dependency chain.
  * witness_limbs      one trace on the calling host thread (stk_trace_generate, Montgomery
                       arithmetic in C++), in the ABI's element layout;
  * witness_device     traces generated ON THE DEVICE (stk_trace_generate_dev): many independent
                       traces in parallel, and ONE trace in parallel chunks when the step
                       polynomials have degree <= 1 (chunk starts from powers of the companion
                       matrix); a single non-linear trace falls back to the host recurrence
                       with its upload overlapped (stk_trace_generate_upload).
Both produce what STARK.mk_proof accepts directly.
The AIR class itself (asserts steps == 511, sympy) is out of scope."""
import numpy as np

from .engine import default_engine
from .limbs import ints_to_limbs, limbs_to_ints
from .modp import element_to_int
from .polynomial import monomials_of


def _monomial_arrays(step_polys, width, p):
  mono_out, mono_coef, mono_exp = [], [], []
  for j, sp in enumerate(step_polys):
    for exps, c in monomials_of(sp, width, p):
      mono_out.append(j)
      mono_coef.append(c)
      mono_exp.append(list(exps))
  nm = len(mono_out)
  h_out = np.asarray(mono_out, dtype=np.uint32)
  h_coef = ints_to_limbs(mono_coef) if nm else np.zeros((0, 8), np.uint32)
  h_exp = np.asarray(mono_exp, dtype=np.uint8).reshape(nm, width) if nm else np.zeros((0, width), np.uint8)
  return nm, h_out, h_coef, h_exp


def witness_limbs(field, inp, steps, width, step_polys, engine=None, out=None):
  """witness[dim][step] as a (width, steps, 8) uint32 array: what STARK.mk_proof accepts
  directly.  `out` may be a preallocated array, e.g. Engine.pinned((width, steps, 8)).array so
  that the prover's upload is one DMA from pinned memory."""
  eng = engine or default_engine()
  p = field.p
  eng.set_field(p)
  assert len(inp) == width
  nm, h_out, h_coef, h_exp = _monomial_arrays(step_polys, width, p)
  h_inp = ints_to_limbs([element_to_int(v) % p for v in inp])
  if out is None:
    out = np.empty((width, steps, 8), dtype=np.uint32)
  assert out.shape == (width, steps, 8) and out.dtype == np.uint32 and out.flags["C_CONTIGUOUS"]
  eng._check(eng.lib.stk_trace_generate(eng.ctx, h_inp.ctypes.data, steps, width, h_out.ctypes.data,
                                        h_coef.ctypes.data, h_exp.ctypes.data, nm, out.ctypes.data))
  return out


def is_affine(step_polys, width, p):
  """True when every step polynomial has total degree <= 1."""
  return all(sum(exps) <= 1 for sp in step_polys for exps, _ in monomials_of(sp, width, p))


def witness_device(field, inp, steps, width, step_polys, engine=None, ntraces=1, pinned=None):
  """Trace(s) generated into device memory: returns a DevBuf holding
  witness[(t*width + dim)*steps + step] (ntraces == 1: the (width, steps) witness mk_proof takes).
  `inp` is one input state, or a list of `ntraces` input states.
  One non-linear trace cannot be split: it is generated on the host into `pinned` (a PinnedBuf
  of shape (width, steps, 8), allocated here when missing) while the finished blocks upload."""
  eng = engine or default_engine()
  p = field.p
  eng.set_field(p)
  nm, h_out, h_coef, h_exp = _monomial_arrays(step_polys, width, p)
  states = [inp] if ntraces == 1 and not isinstance(inp[0], (list, tuple)) else list(inp)
  assert len(states) == ntraces and all(len(s) == width for s in states)
  h_inp = ints_to_limbs([element_to_int(v) % p for s in states for v in s])
  d_w = eng.alloc(ntraces * width * steps * 32)
  if ntraces > 1 or is_affine(step_polys, width, p):
    eng._check(eng.lib.stk_trace_generate_dev(eng.ctx, h_inp.ctypes.data, ntraces, steps, width, h_out.ctypes.data,
                                              h_coef.ctypes.data, h_exp.ctypes.data, nm, d_w.ptr, steps))
  else:
    own = pinned is None
    if own:
      pinned = eng.pinned((width, steps, 8))
    eng._check(eng.lib.stk_trace_generate_upload(eng.ctx, h_inp.ctypes.data, steps, width, h_out.ctypes.data,
                                                 h_coef.ctypes.data, h_exp.ctypes.data, nm, pinned.array.ctypes.data,
                                                 d_w.ptr, steps))
    if own:
      eng.sync()
      pinned.free()
  return d_w


def get_computational_trace(inp, steps, width, step_polys, field=None, engine=None):
  """starks/air.py:31-52: (computational_trace, output) with computational_trace[step][dim]
  as field elements (ints when no field is given... the reference's states are whatever
  step_polys returns; here they are instances of `field`)."""
  if field is None:
    field = type(inp[0])
  w = witness_limbs(field, inp, steps, width, step_polys, engine=engine)
  cols = [limbs_to_ints(w[j]) for j in range(width)]
  trace = [[field(cols[j][i]) for j in range(width)] for i in range(steps)]
  return trace, trace[-1]


def generate_witness(trace):
  """AIR.generate_witness (starks/air.py:124): witness[dim][step]."""
  width = len(trace[0])
  return [[state[j] for state in trace] for j in range(width)]
