"""Measurement aid (not product code): write-only and copy bandwidth of the device with plain torch kernels, to put the store
stream of the curvature kernels (13.2 GB per pass) into perspective.  Prints GB/s."""
import torch

x = torch.empty(1 << 30, dtype=torch.float64, device="cuda")          # 8 GiB
y = torch.empty_like(x)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


t = timed(lambda: x.fill_(1.5))
print("fill  (write only)        %.0f GB/s" % (x.numel() * 8 / t / 1e9))
t = timed(lambda: y.copy_(x))
print("copy  (read + write)      %.0f GB/s total" % (2 * x.numel() * 8 / t / 1e9))
t = timed(lambda: torch.add(x, 1.0, out=y))
print("y = x + 1 (read + write)  %.0f GB/s total" % (2 * x.numel() * 8 / t / 1e9))
z = torch.empty(4, 1 << 28, dtype=torch.float64, device="cuda")       # 1 read : 4 writes, the mix of the gradient / flame-normal passes
src = x[: 1 << 28]
t = timed(lambda: z.copy_(src.expand(4, -1)))
print("1 read : 4 writes         %.0f GB/s total (%.0f written)" % (5 * src.numel() * 8 / t / 1e9, 4 * src.numel() * 8 / t / 1e9))
