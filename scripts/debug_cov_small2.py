import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
exec(open(os.path.join(os.path.dirname(__file__), "debug_cov_small.py")).read().split("g = torch.Generator")[0])
for n in (1600, 20000):
    for c in (8, 23, 31, 32, 33, 40, 48, 49, 56, 63, 64):
        g2 = torch.Generator().manual_seed(c)
        f = torch.relu(torch.randn(1, n, 1, c, generator=g2)).cuda()
        s = torch.relu(1.5 * torch.randn(1, n - 16, 1, c, generator=g2) + 0.25).cuda()
        out = ob.ot_loop(f, s, "pca", 1)
        ref = ref_step(f, s, "pca").float()
        d = (out - ref).abs()
        print(f"n={n} c={c}: err {float(d.max()) / float(ref.abs().max()):.2e}  worst channel {int(d.amax((0, 1, 2)).argmax())} "
              f"mean-out err {float((out.mean((0,1,2)) - ref.mean((0,1,2))).abs().max()):.2e}", flush=True)
