"""Narrow covariance path (cov_small.cu; c <= 64) against an FP64 `eigh` reference computed on the device: one step
and a 3-iteration loop of pca / sym at several shapes and feature scales (|mu| up to 1000, sigma up to 3000), then the
padding check: c = 49 (15 internal padding columns) against c = 64 with 15 explicit zero channels - bit-identical.
OPTEX_COV_SMALL=0 runs the same shapes through the wide path."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import optimaltextures_b200 as ob


def ref_step(p, s, mode, eps=1.0):
    c = p.shape[-1]
    P, S = p.reshape(-1, c).double(), s.reshape(-1, c).double()
    mp, ms = P.mean(0), S.mean(0)
    ct = (P - mp).T @ (P - mp) / P.shape[0] + eps * torch.eye(c, device=p.device, dtype=torch.float64)
    cs = (S - ms).T @ (S - ms) / S.shape[0] + eps * torch.eye(c, device=p.device, dtype=torch.float64)

    def sq(a):
        w, v = torch.linalg.eigh(a)
        return v @ torch.diag(w.clamp_min(0).sqrt()) @ v.T

    qt = sq(ct)
    if mode == "pca":
        T = sq(cs) @ torch.linalg.inv(qt)
    else:
        qi = torch.linalg.inv(qt)
        T = qi @ sq(qt @ cs @ qt) @ qi
    return ((P - mp) @ T.T + ms).reshape(p.shape)


g = torch.Generator(device="cuda").manual_seed(0)
for n, ns, c, scale, off in ((262144, 376832, 23, 1.0, 0.3), (65536, 65536, 32, 1.0, 0.3), (4096, 4096, 64, 1.0, 0.0),
                             (256, 300, 16, 1.0, 0.0), (65536, 65536, 32, 300.0, 1000.0), (100, 90, 3, 1.0, 0.5),
                             (65536, 65536, 32, 3000.0, 0.0)):
    sig = torch.logspace(1.0, -0.5, c, device="cuda") * scale
    p = torch.randn(1, n, 1, c, device="cuda", generator=g) * sig + off
    s = (1.2 * torch.randn(1, ns, 1, c, device="cuda", generator=g) + 0.1) * sig + off
    if c == 32:                          # zero channels like OptimalTexture's padding
        p[..., 23:] = 0
        s[..., 23:] = 0
    for mode in ("pca", "sym"):
        out = ob.optimal_transport(p, s, mode, rotation=torch.eye(c, device="cuda"))
        ref = ref_step(p, s, mode)
        sc = max(1.0, float(ref.abs().max()))
        err = float((out.double() - ref).abs().max()) / sc
        loop = ob.ot_loop(p, s, mode, 3)
        r3 = p
        for _ in range(3):
            r3 = ref_step(r3, s, mode).float()
        e3 = float((loop.double() - r3.double()).abs().max()) / sc
        print(f"n={n} c={c} scale={scale} off={off} {mode}: step err {err:.2e} nan={int(torch.isnan(out).sum())}  "
              f"loop(3) err {e3:.2e} nan={int(torch.isnan(loop).sum())}", flush=True)

# the padding test's shape: c = 49 (CP = 64 with 15 internal padding columns) against c = 64 with 15 explicit zero channels
g2 = torch.Generator().manual_seed(9)
c = 49
f = torch.relu(torch.randn(1, 40, 40, c, generator=g2)).cuda()
s = torch.relu(1.5 * torch.randn(1, 36, 44, c, generator=g2) + 0.25).cuda()
content = torch.relu(torch.randn(1, 40, 40, c, generator=g2)).cuda()


def widen(t):
    w = torch.zeros(*t.shape[:-1], 64, device="cuda")
    w[..., :c] = t
    return w


for mode in ("pca", "sym"):
    a = ob.ot_loop(f, s, mode, 4, content=content, content_strength=0.05)
    b = ob.ot_loop(widen(f), widen(s), mode, 4, content=widen(content), content_strength=0.05)[..., :c]
    r = f
    for _ in range(4):
        r = ref_step(r, s, mode).float()
        r = r + 0.05 * (content - r)
    sc = float(r.abs().max())
    print(f"c=49 {mode}: unpadded vs ref {float((a - r).abs().max()) / sc:.2e}, padded vs ref {float((b - r).abs().max()) / sc:.2e}, "
          f"unpadded vs padded {float((a - b).abs().max()) / sc:.2e}  nan {int(torch.isnan(a).sum())} {int(torch.isnan(b).sum())}")
    a1 = ob.ot_loop(f, s, mode, 1)
    b1 = ob.ot_loop(widen(f), widen(s), mode, 1)[..., :c]
    r1 = ref_step(f, s, mode).float()
    print(f"   one step, no content: unpadded vs ref {float((a1 - r1).abs().max()) / sc:.2e}, padded vs ref "
          f"{float((b1 - r1).abs().max()) / sc:.2e}")
