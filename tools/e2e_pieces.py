import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
n = 1 << 24
x = torch.zeros(n, dtype=torch.complex128, device="cuda")
xr = torch.view_as_real(x)
hx = torch.empty((n, 2), dtype=torch.float64).pin_memory()
print("pinned", hx.is_pinned())
y = torch.zeros((1024, n // 512), dtype=torch.complex128, device="cuda")
def T(f, name, k=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); print(name, (time.perf_counter() - t0) / k * 1e3, "ms")
T(lambda: xr.copy_(hx, non_blocking=True), "h2d")
T(lambda: y[:, :64].abs().sum(dim=1).cpu(), "summary")
T(lambda: torch.view_as_real(y[:, :64]).pow(2).sum(dim=(1, 2)).cpu(), "summary2")
