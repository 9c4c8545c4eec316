"""PCIe ceiling next to the e2e number: pinned H2D alone, D2H alone, both at once (two streams),
with and without binding the process to the GPU's NUMA-local CPUs before the pinned allocation."""
import os, sys, time
sys.path.insert(0, '.')
import torch
def bw(label):
    nbytes = 1 << 30
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    d_b = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / (time.perf_counter() - t0) / 1e9
    run(True, True, 1)
    print('%-28s H2D %.1f GB/s  D2H %.1f GB/s  both %.1f GB/s each way' % (label, run(True, False), run(False, True), run(True, True)), flush=True)
print('cpus', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
bw('default affinity')
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
    print('nvml cpu affinity:', len(cpus), 'cpus', sorted(cpus)[:4], '...')
    os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    bw('GPU-local CPUs')
    others = os.sched_getaffinity(0)
except Exception as ex:
    print('nvml affinity unavailable:', ex)
