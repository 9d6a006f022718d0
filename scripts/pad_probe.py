"""cdf / sort OT loops at PCA'd channel counts (not multiples of 32): the SIMT-GEMM path the product takes today against
zero-padding the channels to a multiple of 32 with a block-diagonal rotation diag(R_c, I) (tensor-core path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob

def run(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

iters = 8
for mode in ("cdf",):
    for (n, c) in [(1024, 310), (4096, 346), (16384, 181), (65536, 85), (262144, 23), (1179648, 85), (2359296, 23)]:
        g = torch.Generator().manual_seed(n + c)
        p = torch.randn(1, n, 1, c, generator=g).cuda(); s = (torch.randn(1, n, 1, c, generator=g) * 1.2 + 0.1).cuda()
        rots = ob.random_rotations(c, iters, "cuda", seed=3)
        t_plain, out_plain = run(lambda: ob.ot_loop(p, s, mode, iters, rotations=rots))
        pad = (-c) % 32
        def widen(t):
            w = torch.zeros(*t.shape[:-1], c + pad, device="cuda"); w[..., :c] = t; return w
        eye = torch.eye(c + pad, device="cuda").repeat(iters, 1, 1); eye[:, :c, :c] = rots
        pw, sw = widen(p), widen(s)
        t_pad, out_pad = run(lambda: ob.ot_loop(pw, sw, mode, iters, rotations=eye))
        d = (out_pad[..., :c] - out_plain).abs()
        print(f"{mode} n={n} c={c}: plain {t_plain / iters * 1e3:.0f} us/iter, padded to {c + pad}: {t_pad / iters * 1e3:.0f} us/iter; "
              f"frac |diff| > 1e-3: {float((d > 1e-3).float().mean()):.4f}", flush=True)
