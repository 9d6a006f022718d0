"""Narrow covariance path (cov_small.cu): us per loop iteration against the row count at c = 32 / 64 - the n -> 0 limit
is the chain kernel + launch latencies, the slope the three N x C passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
mode = sys.argv[1] if len(sys.argv) > 1 else "pca"
g = torch.Generator(device="cuda").manual_seed(0)
for c in (32, 64):
    for hw in (16, 64, 128, 256, 512):
        decay = (0.9 ** torch.arange(c, device="cuda")) * 30.0
        p = torch.randn(1, hw, hw, c, device="cuda", generator=g) * decay * 0.7 + 0.3
        s = torch.randn(1, hw, hw, c, device="cuda", generator=g) * decay + 0.5
        ob.ot_loop(p, s, mode, 5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ob.ot_loop(p, s, mode, 40)
        e1.record()
        torch.cuda.synchronize()
        print(f"{mode} c={c} n={hw * hw}: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us per iteration", flush=True)
