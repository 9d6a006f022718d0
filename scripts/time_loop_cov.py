"""optex_ot_loop at the headline shape for the closed-form modes: us per iteration (style side reused)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.relu(torch.randn(1, 128, 128, 512, device="cuda", generator=g))
s = torch.relu(1.3 * torch.randn(1, 128, 128, 512, device="cuda", generator=g) + 0.2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mode in ("pca", "sym", "chol", "cdf", "sort"):
    iters = 40
    ob.ot_loop(p, s, mode, 5)
    torch.cuda.synchronize()
    e0.record()
    ob.ot_loop(p, s, mode, iters)
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per iteration in optex_ot_loop ({iters} iterations)", flush=True)
