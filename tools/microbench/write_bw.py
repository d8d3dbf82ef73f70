"""Write-only and copy bandwidth of the box (torch fill / copy, CUDA events): the roof of a kernel whose traffic is 99 % stores."""
import torch
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for gb in (4, 16, 32):
    n = gb * (1 << 30) // 8
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    ms = t(lambda: x.zero_())
    print(f"fill {gb} GiB: {ms:.3f} ms  {gb * 1.073741824 / ms * 1e3:.0f} GB/s", flush=True)
    if gb <= 16:
        y = torch.empty_like(x)
        ms = t(lambda: y.copy_(x))
        print(f"copy {gb} GiB: {ms:.3f} ms  {2 * gb * 1.073741824 / ms * 1e3:.0f} GB/s (read + write)", flush=True)
        s = torch.zeros((), dtype=torch.float64, device="cuda")
        ms = t(lambda: x.sum())
        print(f"read {gb} GiB: {ms:.3f} ms  {gb * 1.073741824 / ms * 1e3:.0f} GB/s", flush=True)
        del y
    del x
    torch.cuda.empty_cache()
