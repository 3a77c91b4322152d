"""PCIe H2D rate: copy engine (cudaMemcpyAsync from pinned memory) vs what the in-kernel stage-in reaches (45.7 GB/s)."""
import time
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, n // 4, n // 64, 262144):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for o in range(0, n, chunk):
        d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("chunk %9d B: %.2f ms = %.1f GB/s (host issue time %.2f ms)" % (chunk, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6, t_issue * 1e3))
