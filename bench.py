#!/usr/bin/env python
"""bench.py -- headline benchmark: batched forward NTT over the STARK prime field.

Metric (BASELINE.json): NTT Melem/s at N = 2^20.  One "step" = one forward transform of
`--cols` (default 64) columns of 2^20 elements (2 GiB in, 2 GiB out -> larger than L2, no
flush needed).  `value` times the transforms with inputs resident in HBM (CUDA events on
the launching stream); `e2e` times the same step through the public API with pinned HOST
buffers (H2D + transform + D2H inside the timed region).  N > 1 GPUs: one process per GPU
(torchrun), each rank transforms its own columns (column sharding, no data-path collective,
weak scaling); the time is the max over ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the CPU arm: the reference is pure Python and cannot travel to the
GPU box, so the arm runs the C oracle port of fft_1d (oracle/) on all host threads, on a
bounded sample (one 2^20 column per thread and step)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P = 2**256 - 351 * 2**32 + 1
LOGN = 20
N = 1 << LOGN
INT_OPS_PER_BUTTERFLY = 264       # SURVEY.md 8(d), frozen
BYTES_PER_ELEM = 64               # read once + write once
# measured once per change with ncu (profiles/r01b_ncu_ntt_pass_summary.txt): the two passes of
# one step move 4.45 GB and 4.25 GB; each pass reads and writes the whole 2 GiB batch
NCU_TRAFFIC_BYTES_PER_LAUNCH = 4.35e9


def peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    d = json.load(open(path))
    return float(d["hbm_gbs"]), "measured"
  return 6650.0, "fallback"


class ClockSampler(threading.Thread):
  """Samples SM clock / throttle reasons through NVML every ~5 ms while the timed region runs
  (nvidia-smi as a fallback)."""

  def __init__(self, gpu_index):
    super().__init__(daemon=True)
    self.idx, self.samples, self.reasons, self.maxmhz = gpu_index, [], set(), None
    self._stop_evt = threading.Event()
    self.nvml = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nvml = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
      self.maxmhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
    except Exception:
      self.nvml = None

  @staticmethod
  def _physical_index(i):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
      try:
        return int(vis.split(",")[i])
      except Exception:
        return i
    return i

  def _sample_nvml(self):
    n = self.nvml
    self.samples.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
    for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                      ("hw_thermal_slowdown", 0x40)):
      if r & bit:
        self.reasons.add(name)

  def _sample_smi(self):
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
    self.samples.append(float(out[0]))
    self.maxmhz = float(out[1])
    for nm, v in zip(names, out[2:]):
      if v.strip().lower() == "active":
        self.reasons.add(nm)

  def run(self):
    while not self._stop_evt.is_set():
      try:
        if self.nvml is not None:
          self._sample_nvml()
        else:
          self._sample_smi()
      except Exception:
        pass
      self._stop_evt.wait(0.005 if self.nvml is not None else 0.1)

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=10)
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons),
            "samples": len(s), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def pcie_duplex_ceiling(torch, local):
  """Pinned H2D and D2H copies of 512 MiB at once on two streams: GB/s per direction."""
  nbytes = 1 << 29
  dev = "cuda:%d" % local
  h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
  h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
  d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
  d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
  s1, s2 = torch.cuda.Stream(device=local), torch.cuda.Stream(device=local)
  best = 0.0
  for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
      with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
      with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    if rep:
      best = max(best, 2 * nbytes / (time.perf_counter() - t0) / 1e9)
  return best


def synth_columns(cols, n, seed):
  import numpy as np
  rng = np.random.default_rng(seed)
  a = rng.integers(0, 2**32, size=(cols, n, 8), dtype=np.uint64).astype(np.uint32)
  a[:, :, 7] &= 0x7FFFFFFF  # < 2^255 < p: canonical residues
  return a


def extras(eng, torch, stream, local):
  """Second half of BASELINE.json's metric, per GPU: LDE + Merkle commit of 64 trace columns of
  2^18 steps (config 3; trace resident on the device -> root on the host) and the full
  Fibonacci proof at 2^20 steps (config 5 shape on one GPU)."""
  import numpy as np
  out = {}
  # single-column transforms (SURVEY 8d asks for batch = 1 next to batch = 64)
  for logn in (20, 24):
    n1 = 1 << logn
    w1 = pow(7, (P - 1) // n1, P)
    a = torch.randint(0, 2**31 - 1, (n1, 8), dtype=torch.int32, device="cuda:%d" % local)
    b = torch.empty_like(a)
    for _ in range(3):
      eng.ntt(a.data_ptr(), n1, n1, b.data_ptr(), n1, n1, 1, w1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
      e0.record(stream)
      for _ in range(20):
        eng.ntt(a.data_ptr(), n1, n1, b.data_ptr(), n1, n1, 1, w1)
      e1.record(stream)
    torch.cuda.synchronize()
    out["ntt_melem_per_s_2^%d_batch1" % logn] = n1 / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e6
    del a, b
  steps, ext, ncols = 1 << 18, 8, 64
  n = steps * ext
  g2 = pow(7, (P - 1) // n, P)
  d_tr = torch.randint(0, 2**31 - 1, (ncols, steps, 8), dtype=torch.int32, device="cuda:%d" % local)
  d_ev = torch.empty((ncols, n, 8), dtype=torch.int32, device="cuda:%d" % local)
  d_nodes = torch.empty((n, 32), dtype=torch.uint8, device="cuda:%d" % local)
  for _ in range(2):
    eng.lde_commit(d_tr.data_ptr(), steps, steps, ext, ncols, g2, d_ev.data_ptr(), n, d_nodes.data_ptr())
  reps = 5
  t0 = time.perf_counter()
  for _ in range(reps):
    root = eng.lde_commit(d_tr.data_ptr(), steps, steps, ext, ncols, g2, d_ev.data_ptr(), n, d_nodes.data_ptr())
  out["lde_merkle_commit_ms_64x2^18_x8"] = (time.perf_counter() - t0) / reps * 1e3
  # same call with the Merkle bottom level hashed inside the transform's final pass (opt-in)
  os.environ["STK_FUSED_HASH"] = "1"
  eng.lde_commit(d_tr.data_ptr(), steps, steps, ext, ncols, g2, d_ev.data_ptr(), n, d_nodes.data_ptr())
  t0 = time.perf_counter()
  for _ in range(reps):
    root_u = eng.lde_commit(d_tr.data_ptr(), steps, steps, ext, ncols, g2, d_ev.data_ptr(), n, d_nodes.data_ptr())
  out["lde_merkle_commit_ms_fused_leaf_hash"] = (time.perf_counter() - t0) / reps * 1e3
  del os.environ["STK_FUSED_HASH"]
  assert root_u == root, "fused and separate leaf hashing disagree"
  # split: LDE alone / commit alone (CUDA events on the launching stream)
  e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
  with torch.cuda.stream(stream):
    e[0].record(stream)
    eng.lde(d_tr.data_ptr(), steps, steps, ext, ncols, g2, d_ev.data_ptr(), n)
    e[1].record(stream)
    eng.merkle_commit(d_ev.data_ptr(), n, ncols, n, d_nodes.data_ptr(), want_root=False)
    e[2].record(stream)
  torch.cuda.synchronize()
  out["lde_ms"] = e[0].elapsed_time(e[1])
  out["merkle_ms"] = e[1].elapsed_time(e[2])
  compressions = (n // 2) * ncols + (n // 2 - 1)
  out["merkle_gcompress_per_s"] = compressions / (out["merkle_ms"] * 1e-3) / 1e9
  lde_bfly = ncols * ((steps // 2) * 18 + (n // 2) * 21)
  out["lde_gbutterflies_per_s"] = lde_bfly / (out["lde_ms"] * 1e-3) / 1e9
  del d_tr, d_ev, d_nodes
  torch.cuda.empty_cache()
  # full proof, 2^20 steps
  from starks_b200.limbs import ints_to_limbs
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  from starks_b200.air import witness_limbs
  psteps = 1 << 20
  fib = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
  wpin = eng.pinned((2, psteps, 8))            # witness in pinned host memory, like the NTT inputs
  t0 = time.perf_counter()
  witness = witness_limbs(IntegersModP(P), [0, 1], psteps, 2, fib, engine=eng, out=wpin.array)
  out["trace_generate_s_fib_2^20_steps"] = time.perf_counter() - t0
  a, b = 0, 1
  for _ in range(psteps - 1):
    a, b = b, (a + b) % P
  assert ints_to_limbs([a, b]).tolist() == witness[:, -1, :].tolist(), "generated trace differs from the plain recurrence"
  S = STARK(IntegersModP(P), psteps, 8, 2, [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}], engine=eng)
  for _ in range(2):  # warm-up: tables, buffer pool, first-use transients
    S.mk_proof(witness, [(0, 0, 0), (0, 1, 1)])
  t0 = time.perf_counter()
  for _ in range(3):
    proof = S.mk_proof(witness, [(0, 0, 0), (0, 1, 1)])
  out["stark_proof_s_fib_2^20_steps_x8"] = (time.perf_counter() - t0) / 3
  t0 = time.perf_counter()
  assert S.verify_proof(proof, witness, [(0, 0, 0), (0, 1, 1)])
  out["stark_verify_s_fib_2^20_steps_x8"] = time.perf_counter() - t0
  out["stark_proof_fri_layers"] = len(proof[3])
  out["stark_proof_phases_ms"] = {k: round(v, 2) for k, v in S.timings.items() if k.endswith("_ms")}
  return out


def cpu_port_baseline(cols_sample, threads):
  """Times the oracle (C port of starks/fft.py:303-331) on `cols_sample` columns of 2^20."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  w = pow(7, (P - 1) // N, P)
  data = synth_columns(cols_sample, N, 1234)
  t0 = time.perf_counter()
  orc.fft_limbs(P, w, data, N, nthreads=threads)
  dt = time.perf_counter() - t0
  return cols_sample * N / dt / 1e6, dt


def run_reference(args, rank, world):
  if rank != 0:
    return
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  orc.build()
  threads = max(1, min(orc.threads(), os.cpu_count() or 1))
  cols = threads  # one column per thread and step
  w = pow(7, (P - 1) // N, P)
  data = synth_columns(cols, N, 99)
  steps = max(1, min(args.steps, 3))
  for _ in range(min(args.warmup, 1)):
    orc.fft_limbs(P, w, data[:threads], N, nthreads=threads)
  t0 = time.perf_counter()
  for _ in range(steps):
    orc.fft_limbs(P, w, data, N, nthreads=threads)
  dt = (time.perf_counter() - t0) / steps
  val = cols * N / dt / 1e6
  line = {
      "impl": "reference", "metric": "ntt_melem_per_s_2^20", "value": val, "unit": "Melem/s", "n_gpus": args.gpus,
      "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (256-bit modular integers)", "data": "synthetic",
      "config": {"workload": "forward NTT, %d columns x 2^20 (bounded sample of the 64-column step), STARK prime" % cols,
                 "note": "reference is pure Python (fft_1d 2^20 = 73.7 s/column, BASELINE.md); this arm is the C oracle port on all host threads"},
      "cpu_baseline": {"value": val, "unit": "Melem/s", "cores": threads, "kind": "port",
                       "sample": "%d columns x 2^20 per step, %d steps" % (cols, steps)},
      "e2e": {"value": val, "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours")
  ap.add_argument("--cols", type=int, default=64)
  ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-extras", action="store_true", help="skip the LDE+Merkle / full-proof timings")
  args = ap.parse_args()
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if args.impl == "reference":
    run_reference(args, rank, world)
    return
  warmup = max(args.warmup, 3)

  import numpy as np
  import torch
  import torch.distributed as dist
  from starks_b200 import Engine

  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  eng = Engine(local)
  stream = torch.cuda.Stream(device=local)
  eng.set_stream(stream.cuda_stream)
  cols = args.cols
  w = pow(7, (P - 1) // N, P)

  host_in = eng.pinned((cols, N, 8))
  host_out = eng.pinned((cols, N, 8))
  host_in.array[...] = synth_columns(cols, N, 1000 + rank)
  d_in = torch.empty((cols, N, 8), dtype=torch.int32, device="cuda:%d" % local)
  d_out = torch.empty_like(d_in)
  with torch.cuda.stream(stream):
    d_in.copy_(torch.from_numpy(host_in.array.view(np.int32)), non_blocking=True)
  stream.synchronize()

  def step():
    eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(warmup):
    step()
  barrier()
  sampler = ClockSampler(local)
  sampler.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  with torch.cuda.stream(stream):
    e0.record(stream)
    for _ in range(args.steps):
      step()
    e1.record(stream)
  barrier()
  ms = e0.elapsed_time(e1)
  clocks = sampler.stop()
  # e2e through the host-buffer API
  e2e_s, pcie_duplex = None, None
  if not args.no_e2e:
    e2e_steps = max(1, min(args.steps, 5))
    eng.ntt_host(host_in.array, N, w, out=host_out.array)  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      eng.ntt_host(host_in.array, N, w, out=host_out.array)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    barrier()  # all ranks copy at once, like the e2e steps: the host side is shared
    pcie_duplex = pcie_duplex_ceiling(torch, local)
  if world > 1:
    t = torch.tensor([ms, e2e_s or 0.0], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_max = float(t[0]), float(t[1])
    e2e_s = e2e_max if e2e_s is not None else None
  # parity spot check of the timed output (not timed): inverse round trip of one column
  chk = eng.alloc(N * 32)
  eng.ntt(d_out.data_ptr(), N, N, chk.ptr, N, N, 1, w, inverse=True)
  back = chk.download((N, 8))
  assert (back == host_in.array[0]).all(), "timed NTT output failed the inverse round trip"

  if rank == 0:
    ms_step = ms / args.steps
    elems = world * cols * N
    value = elems / (ms_step * 1e-3) / 1e6
    hbm_peak, peak_src = peaks()
    launches_per_step = 2  # 2^20 = two radix-2^10 passes of ntt_pass_kernel
    alg_bytes = BYTES_PER_ELEM * cols * N + 16 * N
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    butterflies = cols * (N // 2) * LOGN
    int_ops = INT_OPS_PER_BUTTERFLY * butterflies
    mb = {}
    try:
      for which, name in ((7, "imad_iadd3_mixed_gops"), (0, "imad_gops"), (6, "butterfly_gops")):
        best = 0.0
        for _ in range(2):
          mms, ops = eng.microbench(which, 4000 if which != 6 else 1000)
          best = max(best, ops / (mms * 1e-3) / 1e9)
        mb[name] = best
    except Exception as ex:  # pragma: no cover
      mb["error"] = str(ex)
    line = {
        "metric": "ntt_melem_per_s_2^20", "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (256-bit modular integers)", "data": "synthetic",
        "config": {"workload": "forward NTT, %d columns x 2^20 per GPU, p = 2^256-351*2^32+1 (BASELINE configs[1] at its headline size)" % cols,
                   "cols_per_gpu": cols, "log2_n": LOGN, "l2_policy": "inputs (2 GiB) exceed L2, no flush",
                   "parallelism": "column-sharded x%d, no collective" % world},
        "clocks": clocks,
        "gpu_launches": args.steps * launches_per_step,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r01b_ncu_ntt_pass_summary.txt",
                     "alg_bytes_per_launch": alg_bytes / launches_per_step, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                     "kernel": "ntt_pass_kernel<StarkField>", "launches_per_step": launches_per_step,
                     "alg_bytes_per_step": alg_bytes,
                     "note": "the kernel is bound by the integer pipes, not HBM (256-bit modular butterflies: ~41 int32 op per byte against a machine balance of ~4.7): int_roofline below is the binding one; DRAM is 12 % busy in ncu"},
        "int_roofline": {"bound": "int32 pipes", "alg_int32_ops_per_step": int_ops,
                         "achieved_gops": int_ops / (ms_step * 1e-3) / 1e9,
                         "peak_gops": mb.get("imad_iadd3_mixed_gops"), "peak_source": "K0 microbenchmark (IMAD+IADD3 dual issue), same run",
                         "frac": (int_ops / (ms_step * 1e-3) / 1e9 / mb["imad_iadd3_mixed_gops"]) if mb.get("imad_iadd3_mixed_gops") else None,
                         "butterflies_per_s_g": butterflies / (ms_step * 1e-3) / 1e9,
                         "in_register_butterfly_peak_g": mb.get("butterfly_gops"), "microbench": mb},
    }
    if e2e_s is not None:
      line["e2e"] = {"value": elems / e2e_s / 1e6, "unit": "Melem/s", "h2d_bytes_per_step": cols * N * 32,
                     "d2h_bytes_per_step": cols * N * 32, "ms_per_step": e2e_s * 1e3,
                     "gb_per_s_each_way": cols * N * 32 / e2e_s / 1e9,
                     "pcie_duplex_plain_copies_gb_per_s_each_way": pcie_duplex,
                     "note": "host-buffer API (stk_ntt_host): three-slot H2D / transform / D2H pipeline; bound by PCIe -- "
                             "beside it, plain pinned 512 MiB copies both ways at once (two streams) in this run"}
    if not args.no_extras:
      try:
        line["extra"] = extras(eng, torch, stream, local)
      except Exception as ex:  # pragma: no cover
        line["extra"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu:
      v, dt = cpu_port_baseline(4, 1)
      line["cpu_baseline"] = {"value": v, "unit": "Melem/s", "cores": 1, "kind": "port",
                              "sample": "4 columns x 2^20, one pass, %.1f s" % dt,
                              "note": "C oracle port of fft_1d; the pure-Python reference measured 0.0142 Melem/s (BASELINE.md)"}
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
