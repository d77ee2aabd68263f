"""Development probe: ways to solve the 60 symmetric 4096 x 4096 float64 eigenproblems of a run_PCA pass on one GPU."""
import time
import torch
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
mats = []
for h in range(8):
    X = torch.rand(6000, 4096, device=dev, generator=g, dtype=torch.float64)
    C = X.t() @ X / 6000
    mats.append(0.5 * (C + C.t()))
A = torch.stack(mats)
def t(fn, name):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{name:40s} {1e3 * (time.perf_counter() - t0) / 8:8.1f} ms per matrix   top {float(r[0].max()):.6f}", flush=True)
t(lambda: [torch.linalg.eigvalsh(m) for m in mats], "loop eigvalsh fp64")
t(lambda: torch.linalg.eigvalsh(A), "batched eigvalsh fp64 [8,4096,4096]")
t(lambda: [torch.linalg.eigvalsh(m.float()) for m in mats], "loop eigvalsh fp32")
t(lambda: torch.linalg.eigvalsh(A.float()), "batched eigvalsh fp32")
for drv in ("gesvd", "gesvdj", "gesvda"):
    try:
        t(lambda: [torch.linalg.svdvals(m, driver=drv) for m in mats[:2]] * 4, f"svdvals fp64 driver={drv}")
    except Exception as e:  # noqa: BLE001
        print(drv, "failed", type(e).__name__)
# accuracy of the gesvda route on a covariance with a fast-decaying spectrum (like attention maps) and exact low rank
for name, M in (("dense", mats[0]),):
    ref = torch.linalg.eigvalsh(M).flip(0)
    got = torch.linalg.svdvals(M, driver="gesvda")
    print(name, "max |diff| / top:", float((got - ref).abs().max() / ref[0]), " rel err top-64:", float(((got - ref).abs() / ref)[:64].max()))
U = torch.linalg.qr(torch.randn(4096, 4096, device=dev, dtype=torch.float64, generator=g)).Q
lam = torch.exp(-torch.arange(4096, device=dev, dtype=torch.float64) / 40.0)
lam[2000:] = 0
M = (U * lam) @ U.t(); M = 0.5 * (M + M.t())
ref = torch.linalg.eigvalsh(M).flip(0).clamp_min(0)
got = torch.linalg.svdvals(M, driver="gesvda")
d = (got - ref).abs()
print("decaying: max |diff| / top:", float(d.max() / ref[0]), " rel err where ref > 1e-6 top:", float((d / ref)[ref > 1e-6 * ref[0]].max()),
      " where ref > 1e-10:", float((d / ref)[ref > 1e-10 * ref[0]].max()))
