import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
c = 64
g = torch.Generator().manual_seed(3)
p = torch.relu(torch.randn(1, 64, 64, c, generator=g)).cuda()
s = torch.relu(1.3 * torch.randn(1, 48, 80, c, generator=g) + 0.2).cuda()
def unfused(mode, iters, rots):
    ref = p
    for i in range(iters):
        ref = ob.optimal_transport(ref, s, mode, rotation=rots[i])
    return ref
for iters in (1, 2, 3, 5):
    rots = ob.random_rotations(c, iters, "cuda", seed=21, first_counter=100)
    for mode in ("cdf", "sort"):
        ob.set_gemm_mode("auto"); a = unfused(mode, iters, rots); f = ob.ot_loop(p, s, mode, iters, rotations=rots)
        ob.set_gemm_mode("fp32"); b = unfused(mode, iters, rots)
        ob.set_gemm_mode("auto")
        sc = float(a.abs().max())
        def frac(x, y, tol): return float(((x - y).abs() > tol * sc).float().mean())
        print(f"iters={iters} {mode:4s} fused-vs-unfused: >1e-3: {frac(f,a,1e-3):.4f} >1e-4: {frac(f,a,1e-4):.4f} max {float((f-a).abs().max()):.4f} | "
              f"fp32-vs-3xtf32 unfused: >1e-3: {frac(a,b,1e-3):.4f} >1e-4: {frac(a,b,1e-4):.4f} max {float((a-b).abs().max()):.4f}")
