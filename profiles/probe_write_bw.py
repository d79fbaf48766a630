"""Probe: write-only vs copy HBM bandwidth on this GPU (context for the obs_render roofline: the kernel is a
pure write stream, while MEASURED_PEAKS.json:hbm_gbs is a read+write copy)."""
import torch

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

if __name__ == "__main__":
    nbytes = 2_500_000_000
    x = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
    y = torch.empty_like(x)
    t = timeit(lambda: x.zero_())
    print("memset  (write only)      %.1f GB/s" % (nbytes / t / 1e6))
    t = timeit(lambda: x.fill_(1.5))
    print("fill    (write only)      %.1f GB/s" % (nbytes / t / 1e6))
    t = timeit(lambda: y.copy_(x))
    print("copy    (read + write)    %.1f GB/s (bytes moved = 2x tensor)" % (2 * nbytes / t / 1e6))
    t = timeit(lambda: x.sum())
    print("reduce  (read only)       %.1f GB/s" % (nbytes / t / 1e6))
