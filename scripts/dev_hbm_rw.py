"""dev: HBM bandwidth by direction on this GPU (torch kernels): pure write (fill), pure read (sum), copy (read + write)."""
import torch
n = 2621440 * 256   # floats: 2.68 GB, the forward output of the 71 -> 256 layer of one grid_raw step
x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
b = n * 4 / 1e9
t = timeit(lambda: x.fill_(1.0)); print(f"fill  {t:.3f} ms  {b / t:.2f} TB/s written")
t = timeit(lambda: x.sum()); print(f"sum   {t:.3f} ms  {b / t:.2f} TB/s read")
t = timeit(lambda: y.copy_(x)); print(f"copy  {t:.3f} ms  {2 * b / t:.2f} TB/s read+written")
t = timeit(lambda: torch.mul(x, 2.0, out=y)); print(f"scale {t:.3f} ms  {2 * b / t:.2f} TB/s read+written")
