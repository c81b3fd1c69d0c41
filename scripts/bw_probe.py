"""HBM bandwidth probe with torch ops (context for the roofline): pure write, pure read, copy, and 1:4 / 4:1 mixes."""
import torch
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n / 1e3
N = 1 << 30  # bf16 elements: 2 GiB
x = torch.empty(N, dtype=torch.bfloat16, device="cuda").normal_()
y = torch.empty_like(x)
t = timeit(lambda: y.zero_()); print(f"pure write (zero_): {2*N/t/1e9:.0f} GB/s")
t = timeit(lambda: x.view(torch.int16).max()); print(f"pure read (max):    {2*N/t/1e9:.0f} GB/s")
t = timeit(lambda: y.copy_(x)); print(f"copy (1R:1W):       {4*N/t/1e9:.0f} GB/s")
x4 = x[: N // 4]
t = timeit(lambda: torch.cat([x4, x4, x4, x4], out=y)); print(f"1R:4W-ish (cat):    {(2*N//4*4 + 2*N)/t/1e9:.0f} GB/s algorithmic")
t = timeit(lambda: torch.add(x, x, out=y)); print(f"add (1R+1W):        {4*N/t/1e9:.0f} GB/s")
