"""How many Newton-Schulz iterations the pca loop really needs: time and result of ot_loop under OPTEX_NS_CAP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
cap = os.environ.get("OPTEX_NS_CAP", "24")
iters = 10
res = {}
for (n, c) in [(1024, 320), (4096, 352), (16384, 192), (65536, 96), (262144, 32), (16384, 512)]:
    g = torch.Generator().manual_seed(n + c)
    sig = torch.logspace(1.5, -0.5, c)
    p = (torch.randn(1, n, 1, c, generator=g) * sig + 0.3).cuda(); s = ((torch.randn(1, n, 1, c, generator=g) * 1.2 + 0.1) * sig + 0.3).cuda()
    out = ob.ot_loop(p, s, "pca", iters); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = ob.ot_loop(p, s, "pca", iters)
    e1.record(); torch.cuda.synchronize()
    print(f"cap={cap} n={n} c={c}: {e0.elapsed_time(e1) / 3 / iters * 1e3:.0f} us/iter  checksum {float(out.double().abs().mean()):.9f} finite {bool(torch.isfinite(out).all())}", flush=True)
