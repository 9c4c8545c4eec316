"""GPU edge cases: strided / in-place device transforms, the largest single-GPU sizes,
degenerate AIR shapes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1


@pytest.fixture(scope="module")
def eng():
  from starks_b200 import Engine
  e = Engine(0)
  yield e
  e.close()


def rand(shape, seed):
  rng = np.random.default_rng(seed)
  a = rng.integers(0, 2**32, size=shape + (8,), dtype=np.uint64).astype(np.uint32)
  a[..., 7] &= 0x7FFFFFFF
  return a


@pytest.mark.parametrize("logn", [5, 10, 13])
def test_strided_and_in_place(eng, oracle, logn):
  n, batch, n_in = 1 << logn, 3, (1 << logn) - 5
  w = pow(7, (P - 1) // n, P)
  in_stride, out_stride = n + 7, n + 3
  x = rand((batch, in_stride), logn)
  want = oracle.fft_limbs(P, w, np.ascontiguousarray(x[:, :n_in]), n, nthreads=2)
  d_in = eng.alloc(x.nbytes).upload(x)
  d_out = eng.alloc(batch * out_stride * 32)
  eng.ntt(d_in.ptr, n_in, in_stride, d_out.ptr, out_stride, n, batch, w)
  got = d_out.download((batch, out_stride, 8))
  assert (got[:, :n] == want).all()
  # in place (same buffer, full length)
  y = rand((batch, n), logn + 50)
  d = eng.alloc(y.nbytes).upload(y)
  eng.ntt(d.ptr, n, n, d.ptr, n, n, batch, w)
  assert (d.download((batch, n, 8)) == oracle.fft_limbs(P, w, y, n, nthreads=2)).all()
  eng.ntt(d.ptr, n, n, d.ptr, n, n, batch, w, inverse=True)
  assert (d.download((batch, n, 8)) == y).all()


def test_bad_roots_are_rejected(eng):
  d = eng.alloc(64 * 32)
  with pytest.raises(ValueError):   # 7 is not a 64th root of unity
    eng.ntt(d.ptr, 64, 64, d.ptr, 64, 64, 1, 7)
  w128 = pow(7, (P - 1) // 128, P)
  with pytest.raises(ValueError):   # order is 128, not 64
    eng.ntt(d.ptr, 64, 64, d.ptr, 64, 64, 1, w128)
  with pytest.raises(ValueError):   # order 32 < 64
    eng.ntt(d.ptr, 64, 64, d.ptr, 64, 64, 1, pow(w128, 4, P))


def test_largest_single_gpu_transform(eng):
  """2^26 points (2 GiB): inverse(forward(x)) == x and spot values against Horner on a
  sparse polynomial."""
  logn = 26
  n = 1 << logn
  w = pow(7, (P - 1) // n, P)
  import torch
  x = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda")
  y = torch.empty_like(x)
  z = torch.empty_like(x)
  torch.cuda.synchronize()
  eng.ntt(x.data_ptr(), n, n, y.data_ptr(), n, n, 1, w)
  eng.ntt(y.data_ptr(), n, n, z.data_ptr(), n, n, 1, w, inverse=True)
  eng.sync()
  assert torch.equal(x, z)
  sp = torch.zeros((n, 8), dtype=torch.int32, device="cuda")
  idx = [0, 3, n // 2 + 7, n - 1]
  coef = [5, 2**40 + 3, 77, 2**62 + 1]
  from starks_b200.limbs import int_to_limbs, limbs_to_ints
  for i, c in zip(idx, coef):
    sp[i] = torch.from_numpy(int_to_limbs(c).view(np.int32))
  torch.cuda.synchronize()
  eng.ntt(sp.data_ptr(), n, n, y.data_ptr(), n, n, 1, w)
  eng.sync()
  for k in (0, 1, n // 2, n - 1, 987654321 % n):
    got = limbs_to_ints(y[k].cpu().numpy().view(np.uint32).reshape(1, 8))[0]
    xk = pow(w, k, P)
    assert got == sum(c * pow(xk, i, P) for i, c in zip(idx, coef)) % P


def test_width_one_and_constant_terms(eng, oracle):
  """AIRs the Fibonacci goldens do not cover: a single column with a constant term, and a
  step function that ignores one column."""
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  F = IntegersModP(P)
  for width, sp, inp in ((1, [{(2,): 1, (0,): 7}], [3]),
                         (2, [{(0, 1): 1}, {(0, 3): 2, (0, 0): P - 1}], [4, 9])):
    steps = 32
    witness = oracle.computational_trace(P, inp, steps, sp)
    boundary = [(0, j, inp[j]) for j in range(width)]
    want = oracle.StarkOracle(steps, 8, width, sp).mk_proof(witness, boundary)
    S = STARK(F, steps, 8, width, sp, engine=eng)
    got = S.mk_proof(witness, boundary)
    assert got == want
    assert S.verify_proof(got, witness, boundary)


def test_single_column_tree_2p23(eng):
  """The l-tree of the 2^20-step proof: 2^23 single-value leaves; a branch re-hashed on the
  host must give the device root."""
  from starks_b200.merkle_tree import verify_branch
  from starks_b200.limbs import limbs_to_be_bytes
  n = 1 << 23
  import torch
  x = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda")
  nodes = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
  torch.cuda.synchronize()
  root = eng.merkle_commit(x.data_ptr(), n, 1, n, nodes.data_ptr())
  idx = [0, 1, n // 4, n // 2 + 1, n - 1, 5000001]
  for i, br in zip(idx, eng.merkle_paths(x.data_ptr(), n, 1, n, nodes.data_ptr(), idx)):
    assert len(br) == 24
    leaf = verify_branch(root, i, br)
    assert leaf == limbs_to_be_bytes(x[i].cpu().numpy().view(np.uint32).reshape(1, 8)).tobytes()


def test_plain_c_client_of_the_abi(tmp_path):
  """examples/abi_roundtrip.c: a C program that includes only include/starks_b200.h and links
  libstarks_b200.so (no Python, no CUDA headers) -- the boundary really is a C ABI."""
  import os
  import subprocess
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  exe = str(tmp_path / "abi_roundtrip")
  subprocess.check_call(["/usr/bin/gcc", "-O2", "-I" + os.path.join(root, "include"),
                         os.path.join(root, "examples", "abi_roundtrip.c"), "-o", exe,
                         "-L" + os.path.join(root, "starks_b200"), "-lstarks_b200",
                         "-Wl,-rpath," + os.path.join(root, "starks_b200")])
  out = subprocess.run([exe], capture_output=True, text=True)
  assert out.returncode == 0, out.stderr
  assert out.stdout.startswith("abi roundtrip ok, root=")


def test_direct_dft_cross_checks_the_fast_path(eng, oracle):
  """stk_dft_generic (the reference's _simple_ft, any order) against the radix-8 passes and the
  oracle, on power-of-two and non-power-of-two orders."""
  for n in (8, 64, 512):
    w = pow(7, (P - 1) // n, P)
    x = rand((2, n), n)
    d_in, d_a, d_b = eng.alloc(x.nbytes).upload(x), eng.alloc(x.nbytes), eng.alloc(x.nbytes)
    for inv in (False, True):
      eng.ntt(d_in.ptr, n, n, d_a.ptr, n, n, 2, w, inverse=inv)
      eng.dft_generic(d_in.ptr, n, n, d_b.ptr, n, n, 2, w, inverse=inv)
      a, b = d_a.download((2, n, 8)), d_b.download((2, n, 8))
      assert (a == b).all() and (a == oracle.fft_limbs(P, w, x, n, inv=inv)).all()
  eng.set_field(31)
  x = oracle.to_limbs([7, 0, 30, 4, 11, 2]).reshape(1, 6, 8)
  d_in, d_o = eng.alloc(x.nbytes).upload(x), eng.alloc(x.nbytes)
  eng.dft_generic(d_in.ptr, 6, 6, d_o.ptr, 6, 6, 1, pow(3, 5, 31))
  assert oracle.from_limbs(d_o.download((6, 8))) == oracle.fft_1d(31, [7, 0, 30, 4, 11, 2], pow(3, 5, 31))
  eng.set_field(P)


def test_more_columns_than_a_grid_dimension(eng):
  """A multi-pass transform puts the column index in gridDim.y (<= 65535): wider batches run as
  consecutive column groups.  66 000 columns of 2^12: sampled columns equal their own
  single-column transform, forward and inverse."""
  import torch
  n, batch = 1 << 12, 66000
  w = pow(7, (P - 1) // n, P)
  gen = torch.Generator(device="cuda")
  gen.manual_seed(5)
  x = torch.randint(0, 2**31 - 1, (batch, n, 8), dtype=torch.int32, device="cuda", generator=gen)
  y = torch.empty_like(x)
  torch.cuda.synchronize()
  for inverse in (False, True):
    eng.ntt(x.data_ptr(), n, n, y.data_ptr(), n, n, batch, w, inverse=inverse)
    eng.sync()
    one = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    for c in (0, 1, 32767, 32768, 65535, 65536, batch - 1):
      eng.ntt(x[c].data_ptr(), n, n, one.data_ptr(), n, n, 1, w, inverse=inverse)
      eng.sync()
      assert torch.equal(one, y[c]), (inverse, c)
  del x, y
  torch.cuda.empty_cache()
