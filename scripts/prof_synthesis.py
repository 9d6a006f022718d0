"""A short synthesis (OptimalTexture.forward) for ncu launch lists: size / iters / passes / mode from argv."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import optimaltextures_b200 as ob
from optimaltextures_b200 import texture, vgg

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = sys.argv[4] if len(sys.argv) > 4 else "pca"
g = torch.Generator().manual_seed(0)
style = torch.rand(1, 3, round(size * 736 / 512 / 32) * 32, size, generator=g).cuda()
pastiche = torch.rand(1, 3, size, size, generator=g).cuda()
model = texture.OptimalTexture(size=size, iters=iters, passes=passes, hist_mode=mode,
                               state_dicts=vgg.random_state_dicts(0))
ob.manual_seed(0)
model.profile = {}
out = model.forward(pastiche, [style])
torch.cuda.synchronize()
print("done", tuple(out.shape), model.ot_calls, {k: round(v, 2) for k, v in model.stage_ms().items()})
