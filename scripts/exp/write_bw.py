"""Probe: pure-write HBM bandwidth on the B200 (fill) next to copy bandwidth, to interpret the fused forward's 658 MB of stores."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 28                       # 1 GiB of fp32
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.0)); print("fill  1 GiB: %.3f ms  %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero  1 GiB: %.3f ms  %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy  1 GiB: %.3f ms  %.0f GB/s read+written" % (ms, 2 * n * 4 / ms / 1e6))
ms = t(lambda: torch.sum(x)); print("sum   1 GiB: %.3f ms  %.0f GB/s read" % (ms, n * 4 / ms / 1e6))
