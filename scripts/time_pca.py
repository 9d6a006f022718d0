"""Time fit_pca / pca_project at the layer shapes of a 1024^2 image (CUDA events, after warm-up)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, stream_ptr, workspace

dev = torch.device("cuda")
lib = _lib.lib()
for n, c in ((368, 512), (1472, 512), (4096, 512), (16384, 512), (65536, 256), (262144, 128), (1048576, 64)):
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn(n, c, generator=g) @ (torch.randn(c, c, generator=g) * (2.0 / c ** 0.5)) + 0.3).to(dev)
    vec = torch.empty(c, c, device=dev)
    sig = torch.empty(c, device=dev)
    kd = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = workspace(dev, lib.optex_fit_pca_workspace_bytes(n, c))
    st = stream_ptr(dev)

    def run():
        call("optex_fit_pca_warm", ptr(x), n, c, ptr(vec), ptr(sig), ptr(kd), None, 0, C.c_void_p(kd.data_ptr() + 4),
             ptr(ws), ws.numel(), st)

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    k, sweeps = kd.tolist()
    xc = (x - x.mean()).double()
    gd = xc.T @ xc
    vd = vec.double()
    d = vd.T @ gd @ vd
    off = float((d - torch.diag(torch.diag(d))).abs().max() / d.diag().max())
    orth = float((vd.T @ vd - torch.eye(c, device=dev, dtype=torch.float64)).abs().max())
    lam = torch.linalg.eigvalsh(gd).flip(0).clamp_min(0).sqrt()
    dsig = float((sig.double() - lam).abs().max() / lam[0])
    v = vec[:, :max(k, 32) // 32 * 32].contiguous()
    for _ in range(2):
        f = ob.pca_project(x, v)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(5):
        f = ob.pca_project(x, v)
    p1.record()
    torch.cuda.synchronize()
    print(f"fit_pca n={n} c={c}: {e0.elapsed_time(e1) / 3:.2f} ms  sweeps={sweeps}  offdiag={off:.1e} orth={orth:.1e} dsigma={dsig:.1e}  k={k}  sigma[0]={float(sig[0]):.3f}  "
          f"project k={v.shape[1]}: {p0.elapsed_time(p1) / 5 * 1e3:.1f} us", flush=True)
