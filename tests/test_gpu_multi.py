"""Multi-GPU parity under pytest: launches tests/dist_worker.py with torchrun over 2 (and 4, 8
when the box has them) GPUs.  Needs at least two devices -- ranks that wait on one another must
never share a GPU -- so it is skipped on the one-GPU test box and run with `gpurun --gpus N`.
The CPU side of the same plumbing (index maps, exchanges over gloo) is tests/test_dist_gloo.py."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ngpus():
  try:
    import torch
    return torch.cuda.device_count()
  except Exception:
    return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_paths_equal_one_gpu(world):
  if _ngpus() < world:
    pytest.skip("needs %d GPUs on this box" % world)
  full = os.environ.get("STK_DIST_FULL", "1") != "0"
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
         "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
  if full:
    cmd.append("--full")
  p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
  tail = (p.stdout[-3000:] + "\n---- stderr ----\n" + p.stderr[-3000:])
  assert p.returncode == 0, tail
  lines = [l for l in p.stdout.splitlines() if l.startswith("DIST_WORKER_RESULT ")]
  assert lines, tail
  res = json.loads(lines[-1][len("DIST_WORKER_RESULT "):])
  assert res["world"] == world and res["dist_ntt_nccl_ok"] and res["dist_ntt_p2p_2^20_ok"]
  assert res.get("sharded_commit_8x2^12_ok") or 8 % world
  if full:
    assert res["dist_ntt_p2p_2^26_ok"] and res["sharded_commit_64x2^18_ok"]
  print(res)
