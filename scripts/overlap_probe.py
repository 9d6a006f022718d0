"""Next-round experiment (NOT yet run on a B200): OptimalTexture(overlap_style=True) against the default schedule on
the bench's 512^2 synthesis - same output (bit for bit is expected: the kernels and their arguments are the same, only
the stream they are enqueued on differs) and the wall-clock time of both.  Usage: python scripts/overlap_probe.py [size]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import optimaltextures_b200 as ob
from optimaltextures_b200 import texture, vgg

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sd = vgg.random_state_dicts(0)
g = torch.Generator().manual_seed(0)
style = torch.rand(1, 3, round(size * 736 / 512 / 32) * 32, size, generator=g).cuda()
pastiche = torch.rand(1, 3, size, size, generator=g).cuda()
results = {}
for overlap in (False, True):
    model = texture.OptimalTexture(size=size, iters=500, passes=5, hist_mode="pca", state_dicts=sd,
                                   overlap_style=overlap)
    for rep in range(3):
        ob.manual_seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model.forward(pastiche, [style])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    results[overlap] = out
    print(f"overlap_style={overlap}: {dt * 1e3:.1f} ms, finite={bool(torch.isfinite(out).all())}, "
          f"sweeps={model.pca_sweeps}", flush=True)
a, b = results[False], results[True]
print(f"max |overlap - default| = {float((a - b).abs().max()):.3e} (output scale {float(a.abs().max()):.3f}), "
      f"identical={torch.equal(a, b)}")
