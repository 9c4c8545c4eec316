import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'oracle')
import numpy as np, torch, torch.distributed as dist
import oracle as orc
from starks_b200 import Engine
from starks_b200 import dist as sd
P = orc.P_STARK
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
for logn in (8, 9, 10):
  n = 1 << logn; L = n // world
  w = pow(7, (P - 1) // n, P)
  x = [orc.synth(1, i) for i in range(n)]
  y = [x[rank + world * m] for m in range(L)]
  H = n // 2
  while H >= world:
    hl = H // world
    for m in range(L):
      if (m // hl) % 2 == 0:
        J = (m * world) | rank
        a, b = y[m], y[m + hl]
        y[m] = (a + b) % P
        y[m + hl] = (a - b) * pow(w, (J % H) * (n // (2 * H)), P) % P
    H //= 2
  t = torch.from_numpy(orc.to_limbs([x[rank + world * m] for m in range(L)]).view(np.int32).copy()).to(dev)
  sd._adopt_stream(eng, t)
  eng.ntt_dist_phase(0, t.data_ptr(), t.data_ptr(), L, 1, L, w, world, rank, False)
  torch.cuda.synchronize()
  got = orc.from_limbs(t.cpu().numpy().view(np.uint32))
  bad = [i for i in range(L) if got[i] != y[i]]
  print("rank", rank, "logn", logn, "KMAX", os.environ.get("STK_NTT_KMAX"), "phase0 mismatches:", len(bad), bad[:10], flush=True)
dist.barrier(); dist.destroy_process_group()
