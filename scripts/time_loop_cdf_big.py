"""cdf loop iterations at the padded PCA'd layer shapes of configs[4] (2048 x 1152 colour transfer, last pass):
us per iteration and the HBM bytes an iteration has to move at least (P, S read; P written) against the copy peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
lib = ob._lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
for (h, w, c), (hs, ws) in (((72, 128, 320), (68, 120)), ((144, 256, 352), (136, 240)), ((288, 512, 192), (272, 480)),
                            ((576, 1024, 96), (544, 960)), ((1152, 2048, 32), (1088, 1920))):
    p = torch.randn(1, h, w, c, device="cuda", generator=g)
    s = torch.randn(1, hs, ws, c, device="cuda", generator=g) * 1.2 + 0.1
    it = 6
    ob.ot_loop(p, s, "cdf", 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.optex_launch_count()
    e0.record()
    ob.ot_loop(p, s, "cdf", it)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3
    mb = (2 * p.numel() + s.numel()) * 4 / 1e6
    print(f"cdf n={h * w} n_s={hs * ws} c={c}: {us:8.1f} us per iteration, {(lib.optex_launch_count() - l0) / it:.1f} launches; "
          f"minimum traffic {mb:.0f} MB = {mb / us:.2f} TB/s effective", flush=True)
