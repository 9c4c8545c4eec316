"""Out-of-bounds WRITES of the production kernels, caught with canaries (compute-sanitizer is
closed on this pool): every output buffer of an entry point sits inside a
larger allocation pre-filled with a pattern, gaps between strided columns included, and after
the call every byte the entry point does not own must still hold the pattern.  Edge geometries on
purpose: single-pass / two-pass / three-pass transforms, 2048-element tiles, zero-padded and
coset routes, ragged batches, strides wider than the rows, in-place use, and the P2P scatter
kernels with all "peer" buffers on one GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1
PAT = 0xA5C35A3C
GUARD = 4096          # elements of canary before and after every buffer


@pytest.fixture(scope="module")
def eng():
  from starks_b200 import Engine
  e = Engine(0)
  yield e
  e.close()


def rand(shape, seed):
  rng = np.random.default_rng(seed)
  a = rng.integers(0, 2**32, size=tuple(shape) + (8,), dtype=np.uint64).astype(np.uint32)
  a[..., 7] &= 0x7FFFFFFF
  return a


class Guarded(object):
  """`elems` elements of 32 bytes between two guards, everything pre-filled with the pattern."""

  def __init__(self, eng, elems):
    self.eng, self.elems = eng, elems
    self.buf = eng.alloc((elems + 2 * GUARD) * 32)
    self.buf.upload(np.full(((elems + 2 * GUARD), 8), PAT, dtype=np.uint32))
    self.ptr = self.buf.ptr + GUARD * 32

  def check(self, owned):
    """owned: boolean mask over the `elems` elements the call may have written."""
    got = self.buf.download((self.elems + 2 * GUARD, 8))
    assert (got[:GUARD] == PAT).all(), "write BELOW the buffer"
    assert (got[GUARD + self.elems:] == PAT).all(), "write ABOVE the buffer"
    body = got[GUARD:GUARD + self.elems]
    free = ~np.asarray(owned, dtype=bool)
    assert (body[free] == PAT).all(), "write into a gap the call does not own"
    return body

  def free(self):
    self.buf.free()


def owned_cols(batch, stride, n, total):
  m = np.zeros(total, dtype=bool)
  for b in range(batch):
    m[b * stride:b * stride + n] = True
  return m


@pytest.mark.parametrize("logn,batch,pad", [(3, 1, 0), (5, 7, 3), (10, 5, 0), (11, 3, 9), (12, 3, 1), (16, 2, 5),
                                            (20, 2, 0), (21, 1, 7), (22, 1, 0)])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_stays_inside_its_columns(eng, logn, batch, pad, inverse):
  n = 1 << logn
  w = pow(7, (P - 1) // n, P)
  stride = n + pad
  x = rand((batch, n), logn)
  d_in = eng.alloc(x.nbytes).upload(x)
  out = Guarded(eng, batch * stride)
  eng.ntt(d_in.ptr, n, n, out.ptr, stride, n, batch, w, inverse=inverse)
  body = out.check(owned_cols(batch, stride, n, batch * stride))
  want = eng.ntt_host(x, n, w, inverse=inverse)
  for b in range(batch):
    assert (body[b * stride:b * stride + n] == want[b]).all()
  # in place, strided
  io = Guarded(eng, batch * stride)
  for b in range(batch):
    eng._check(eng.lib.stk_memcpy_h2d(eng.ctx, io.ptr + b * stride * 32, x[b].ctypes.data, n * 32))
  eng.sync()
  eng.ntt(io.ptr, n, stride, io.ptr, stride, n, batch, w, inverse=inverse)
  body = io.check(owned_cols(batch, stride, n, batch * stride))
  for b in range(batch):
    assert (body[b * stride:b * stride + n] == want[b]).all()
  for g in (out, io):
    g.free()
  d_in.free()


@pytest.mark.parametrize("logsteps,ncols,pad", [(3, 3, 2), (7, 5, 0), (8, 2, 11), (11, 9, 1), (15, 3, 0), (18, 2, 5)])
def test_lde_and_commit_stay_inside(eng, logsteps, ncols, pad):
  """Zero-padded / coset routes with the copied residue-0 coset, evaluation rows wider than N,
  coefficient rows wider than steps; the node buffer of the commit."""
  steps, ext = 1 << logsteps, 8
  n = steps * ext
  g2 = pow(7, (P - 1) // n, P)
  es, cs = n + pad, steps + pad
  tr = rand((ncols, steps), logsteps + 50)
  d_tr = eng.alloc(tr.nbytes).upload(tr)
  ev, co, nodes = Guarded(eng, ncols * es), Guarded(eng, ncols * cs), Guarded(eng, n)
  eng.lde(d_tr.ptr, steps, steps, ext, ncols, g2, ev.ptr, es, d_coeffs=co.ptr, coeff_stride=cs)
  body = ev.check(owned_cols(ncols, es, n, ncols * es))
  co.check(owned_cols(ncols, cs, steps, ncols * cs))
  for c in range(ncols):
    assert (body[c * es:c * es + n][::ext] == tr[c]).all()
  eng.merkle_commit(ev.ptr, n, ncols, es, nodes.ptr)
  nodes.check(np.ones(n, dtype=bool))
  ev.check(owned_cols(ncols, es, n, ncols * es))
  for g in (ev, co, nodes):
    g.free()
  d_tr.free()


@pytest.mark.parametrize("world,logsteps,ncols", [(2, 9, 4), (4, 9, 8), (8, 10, 8), (8, 12, 6)])
def test_row_scatter_stays_inside_the_owner_buffers(eng, world, logsteps, ncols):
  """stk_lde_p2p / stk_ntt_p2p: the final pass stores rows into G "peer" buffers of
  (columns x N/G) elements; uneven column splits and a rank without columns included."""
  from starks_b200.dist import split_columns
  steps, ext = 1 << logsteps, 8
  n = steps * ext
  n_local = n // world
  g2 = pow(7, (P - 1) // n, P)
  tr = rand((ncols, steps), world + logsteps)
  d_tr = eng.alloc(tr.nbytes).upload(tr)
  bufs = [Guarded(eng, ncols * n_local) for _ in range(world)]
  d_ev = eng.alloc(ncols * n * 32)
  eng.lde(d_tr.ptr, steps, steps, ext, ncols, g2, d_ev.ptr, n)
  ev = d_ev.download((ncols, n, 8))
  q, lq = n // 4, n // 4 // world
  want = [np.concatenate([ev[:, j * q + r * lq:j * q + (r + 1) * lq] for j in range(4)], axis=1) for r in range(world)]
  splits = split_columns(ncols, world)
  for r, (c0, c1) in enumerate(splits):
    if c1 > c0:
      eng.lde_p2p(d_tr.at(c0 * steps * 32), steps, steps, ext, c1 - c0, g2, world, c0, [b.ptr for b in bufs])
  for r in range(world):
    body = bufs[r].check(np.ones(ncols * n_local, dtype=bool))
    assert (body.reshape(ncols, n_local, 8) == want[r]).all()
  # the same scatter from coefficient rows (the sharded prover's path), into fresh buffers
  for b in bufs:
    b.free()
  bufs = [Guarded(eng, ncols * n_local) for _ in range(world)]
  d_co = eng.alloc(ncols * steps * 32)
  eng.ntt(d_tr.ptr, steps, steps, d_co.ptr, steps, steps, ncols, pow(g2, ext, P), inverse=True)
  for r, (c0, c1) in enumerate(splits):
    eng.ntt_p2p(d_co.at(c0 * steps * 32), steps, steps, n, c1 - c0, g2, world, c0, [b.ptr for b in bufs])
  for r in range(world):
    body = bufs[r].check(np.ones(ncols * n_local, dtype=bool))
    assert (body.reshape(ncols, n_local, 8) == want[r]).all()
  for b in bufs:
    b.free()
  for b in (d_tr, d_ev, d_co):
    b.free()


@pytest.mark.parametrize("world,logn", [(2, 9), (4, 12), (8, 15), (8, 20)])
def test_four_step_exchange_stays_inside(eng, world, logn):
  n = 1 << logn
  L = n // world
  w = pow(7, (P - 1) // n, P)
  x = rand((n,), logn + 7)
  recv = [Guarded(eng, L) for _ in range(world)]
  work = Guarded(eng, L)
  for r in range(world):
    eng._check(eng.lib.stk_memcpy_h2d(eng.ctx, work.ptr, np.ascontiguousarray(x[r::world]).ctypes.data, L * 32))
    eng.sync()
    eng.ntt_dist_phase0_p2p(work.ptr, L, w, world, r, [b.ptr for b in recv])
  for b in recv + [work]:
    b.check(np.ones(L, dtype=bool))
    b.free()


def test_fold_rows_traces_and_paths_stay_inside(eng):
  from starks_b200.air import _monomial_arrays
  from starks_b200.limbs import ints_to_limbs
  n, world = 1 << 14, 4
  q, lq = n // 4, n // 16
  w = pow(7, (P - 1) // n, P)
  vals = rand((n,), 3)
  for r in range(world):
    rows = np.concatenate([vals[j * q + r * lq:j * q + (r + 1) * lq] for j in range(4)])
    d_rows = eng.alloc(rows.nbytes).upload(np.ascontiguousarray(rows))
    out = Guarded(eng, lq)
    eng.fri_fold4_rows(d_rows.ptr, n, w, 0x1234567, lq, r * lq, out.ptr)
    out.check(np.ones(lq, dtype=bool))
    out.free()
    d_rows.free()
  # device traces with a stride wider than the trace, ragged last chunk
  sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
  nm, h_out, h_coef, h_exp = _monomial_arrays(sp, 2, P)
  for steps, ntr in ((5000, 1), (4097, 3), (300, 5)):
    stride = steps + 13
    out = Guarded(eng, ntr * 2 * stride)
    h_inp = ints_to_limbs([v for t in range(ntr) for v in (t, t + 1)])
    eng._check(eng.lib.stk_trace_generate_dev(eng.ctx, h_inp.ctypes.data, ntr, steps, 2, h_out.ctypes.data,
                                              h_coef.ctypes.data, h_exp.ctypes.data, nm, out.ptr, stride))
    eng.sync()
    out.check(owned_cols(ntr * 2, stride, steps, ntr * 2 * stride))
    out.free()
