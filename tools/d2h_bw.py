import torch, time
n = 1_000_000_000  # 8 GB
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()
def run(k):
    streams = [torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize()
    t0 = time.time()
    c = n // k
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            h[i * c:(i + 1) * c].copy_(d[i * c:(i + 1) * c], non_blocking=True)
    torch.cuda.synchronize()
    return n * 8 / (time.time() - t0) / 1e9
for k in (1, 1, 2, 4, 8):
    print(k, "streams", round(run(k), 1), "GB/s", flush=True)
