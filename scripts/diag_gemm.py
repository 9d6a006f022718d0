"""Diagnostic: rotation GEMM errors per arithmetic mode (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import optimaltextures_b200 as ob
from oracle import rotation as rot_oracle

torch.manual_seed(0)
for (n, c) in [(128, 32), (256, 64), (4096, 64), (1024, 128), (4096, 256), (16384, 512), (4096, 320)]:
    x = torch.relu(torch.randn(n, c)).cuda()
    r = torch.from_numpy(rot_oracle.haar_rotation_qr(c, 1)).float().cuda()
    ref = x.double() @ r.double()
    scale = float(ref.abs().max())
    for mode in ("fp32", "tf32", "tf32x3"):
        ob.set_gemm_mode(mode)
        try:
            xt = ob.rotate_forward(x, r)
            e1 = float((xt.double().T - ref).abs().max()) / scale
            back = ob.rotate_inverse(xt, r)
            ref2 = xt.double().T @ r.double().T
            e2 = float((back.double() - ref2).abs().max()) / float(ref2.abs().max())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                xt = ob.rotate_forward(x, r)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(10):
                back = ob.rotate_inverse(xt, r)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"n={n:6d} c={c:4d} {mode:7s} fwd_err={e1:.2e} inv_err={e2:.2e} fwd={1e5*(t1-t0):8.1f}us inv={1e5*(t2-t1):8.1f}us", flush=True)
        except Exception as ex:
            print(f"n={n} c={c} {mode}: {type(ex).__name__}: {ex}", flush=True)
ob.set_gemm_mode("auto")
