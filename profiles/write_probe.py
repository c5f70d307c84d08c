"""What does a pure write stream reach on this B200?  (denominator sanity for the write-only k_obs)"""
import torch
n = 2_552_000_000 // 4
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def timeit(f, k=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
ms = timeit(lambda: x.fill_(1.0)); print("fill_ 2.55 GB: %.3f ms  %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = timeit(lambda: x.zero_()); print("zero_ 2.55 GB: %.3f ms  %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = timeit(lambda: y.copy_(x)); print("copy  2.55 GB: %.3f ms  %.0f GB/s read+write" % (ms, 2 * n * 4 / ms / 1e6))
