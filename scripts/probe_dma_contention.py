"""Scratch probe: does host->device DMA running in the background slow the (L2-resident) Krylov kernels down?
Device-resident PGD iteration at 1024^2 x M, timed alone and with a continuous H2D copy loop on a side stream."""
import os, sys, time, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
import vch_b200_native as nat
N, M = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 100
c = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4)
rng = np.random.default_rng(42)
phi0 = 0.1 * rng.standard_normal((N + 1, N + 1)); phi0 -= phi0.mean()
dts = np.full(M, 1e-2); t = 1e-2 * np.arange(M + 1); x = np.linspace(0, 1, N + 1)
hist, _, _ = c.forward(torch.from_numpy(phi0).cuda(), None, dts)
xx, yy = np.meshgrid(x, x, indexing="ij")
phiT = torch.from_numpy(0.7 * np.sin(2 * np.pi * xx) * np.cos(np.pi * yy)).cuda()
s_ = torch.from_numpy(t / t[-1]).cuda()[:, None, None]
phiQ = (1 - s_) * hist[0] + s_ * phiT
u = torch.zeros_like(hist)
def run(tag):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p, q, r = c.adjoint(hist, t, 5.0, 10.0, phiQ, phiT, want_pq=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    h2, _, _ = c.forward(hist[0], None, dts)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{tag}: adjoint {1e3*(t1-t0)/M:.3f} ms/level, forward {1e3*(t2-t1)/M:.3f} ms/step", flush=True)
run("warm"); run("alone")
pinned = torch.empty(1 << 27, dtype=torch.float64).pin_memory()      # 1 GiB
dst = torch.empty_like(pinned, device="cuda")
side = torch.cuda.Stream()
stop = False
def pump(direction):
    with torch.cuda.stream(side):
        while not stop:
            if direction == "h2d": dst.copy_(pinned, non_blocking=True)
            else: pinned.copy_(dst, non_blocking=True)
            side.synchronize()
for direction in ("h2d", "d2h"):
    stop = False
    th = threading.Thread(target=pump, args=(direction,)); th.start()
    time.sleep(0.2)
    run(f"with background {direction}")
    stop = True; th.join()
run("alone again")
