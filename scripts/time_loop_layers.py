"""optex_ot_loop per iteration at the PCA'd layer shapes of the last pass of a 512^2 synthesis (configs[1]):
(n, c) = (1024, 320) (4096, 352) (16384, 192) (65536, 96) (262144, 32); blocks with a PCA-like decaying spectrum.
Usage: python scripts/time_loop_layers.py [mode] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import optimaltextures_b200 as ob

mode = sys.argv[1] if len(sys.argv) > 1 else "pca"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lib = ob._lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
total = 0.0
for (hw, c), weight in (((32, 320), 160), ((64, 352), 34), ((128, 192), 52), ((256, 96), 87), ((512, 32), 160)):
    decay = (0.985 ** torch.arange(c, device="cuda")) * 30.0
    p = torch.randn(1, hw, hw, c, device="cuda", generator=g) * decay * 0.7 + 0.3
    s = torch.randn(1, hw, hw, c, device="cuda", generator=g) * decay + 0.5
    ob.ot_loop(p, s, mode, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.optex_launch_count()
    e0.record()
    out = ob.ot_loop(p, s, mode, iters)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    total += us * weight
    print(f"{mode} n={hw * hw} c={c}: {us:.1f} us per iteration, {(lib.optex_launch_count() - l0) / iters:.1f} launches "
          f"per iteration; x{weight} iterations = {us * weight / 1e3:.1f} ms  finite={bool(torch.isfinite(out).all())}",
          flush=True)
print(f"weighted total of the five layers' loops: {total / 1e3:.1f} ms")
