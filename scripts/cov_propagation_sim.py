"""CPU study (numpy) for DESIGN 8-7: inside a pca-mode loop without content blend the pastiche covariance of iteration
i + 1 follows from iteration i, Sigma' = G (Sigma - eps I) G^T + eps I.  How far does the PROPAGATED covariance (fp32
products) drift from the MEASURED one (fp32 Gram of the stored fp32 block), and what does that do to the loop's output?
Reference: the same loop with every covariance measured in float64.  Usage: python scripts/cov_propagation_sim.py [c] [n] [iters]"""
import sys

import numpy as np

c = int(sys.argv[1]) if len(sys.argv) > 1 else 96
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
f32 = np.float32
rng = np.random.default_rng(0)
sig = np.logspace(1.5, -0.5, c)                       # PCA-like spectrum: condition ~1e4 after squaring
S = ((1.2 * rng.standard_normal((n, c)) + 0.1) * sig + 0.3).astype(f32)
P0 = (rng.standard_normal((n, c)) * sig * 0.7 + 0.3).astype(f32)


def sqrtm(a):
    w, v = np.linalg.eigh(a)
    return (v * np.sqrt(np.clip(w, 0, None))) @ v.T


def cov(x, dt):
    x = x.astype(dt)
    xc = x - x.mean(0)
    return (xc.T @ xc / len(x)).astype(dt)


def run(mode):
    p = P0.copy()
    qs = sqrtm(cov(S, np.float64) + np.eye(c))
    mu_s = S.astype(np.float64).mean(0)
    sig_t = None
    for it in range(iters):
        if mode == "f64":
            ct = cov(p, np.float64) + np.eye(c)
        elif mode == "measured" or sig_t is None or (mode.startswith("prop") and "/" in mode and it % int(mode.split("/")[1]) == 0):
            ct = cov(p, f32).astype(np.float64) + np.eye(c)
        else:
            ct = sig_t
        g = qs @ np.linalg.inv(sqrtm(ct))
        g32 = g.astype(f32)
        mu_p = p.astype(np.float64).mean(0)
        p = ((p - mu_p.astype(f32)) @ g32.T + mu_s.astype(f32)).astype(f32)
        if mode.startswith("prop"):
            sig_t = ((g32 @ (ct - np.eye(c)).astype(f32) @ g32.T).astype(f32)).astype(np.float64) + np.eye(c)
    return p


ref = run("f64")
scale = np.abs(ref).max()
for mode in ("measured", "prop", "prop/8"):
    out = run(mode)
    print(f"c={c} n={n} iters={iters} {mode:9s}: max |out - f64 loop| / scale = {np.abs(out - ref).max() / scale:.2e}")
