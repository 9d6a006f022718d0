"""configs[1] (512^2 synthesis, real weights / graffiti.jpg from baseline/_ref) under the schedule options of
OptimalTexture: overlap_style x pca_warm_start.  Wall clock (3rd run), stage split, identity of the outputs.
Usage: python scripts/synth_probe.py [size] [hist_mode]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import optimaltextures_b200 as ob
from baseline import reference
from optimaltextures_b200 import texture

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = sys.argv[2] if len(sys.argv) > 2 else "pca"
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ref = reference.load()
root = ref.path
sd = {}
for d in range(1, 6):
    sd[("encoder", d)] = torch.load(os.path.join(root, "models", f"vgg_normalised_conv{d}_1.pth"), map_location="cpu")
    sd[("decoder", d)] = torch.load(os.path.join(root, "models", f"feature_invertor_conv{d}_1.pth"), map_location="cpu")
styles = ref.util.load_styles([os.path.join(root, "style", "graffiti.jpg")], size=size, scale=1.0)
torch.manual_seed(0)
pastiche = torch.rand(1, 3, size, size).cuda()
dev_styles = [s.cuda() for s in styles]
lib = ob._lib.lib()
outs = {}
for overlap, warm in ((False, False), (True, False), (False, True), (True, True)):
    model = texture.OptimalTexture(size=size, iters=500, passes=passes, hist_mode=mode, state_dicts=sd,
                                   overlap_style=overlap, pca_warm_start=warm)
    for rep in range(3):
        ob.manual_seed(0)
        model.profile = {} if rep == 2 else None
        l0 = lib.optex_launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model.forward(pastiche, dev_styles)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = lib.optex_launch_count() - l0
    outs[(overlap, warm)] = out
    st = {k: round(v, 1) for k, v in model.stage_ms().items()}
    pairs = [round(a.elapsed_time(b), 2) for a, b in model.profile.get("fit_pca", [])]
    print("   fit_pca event pairs (ms):", pairs)
    print(f"overlap={overlap} warm={warm}: {dt * 1e3:.1f} ms  launches={launches}  stages={st}  k={model.last_pca_k} "
          f"sweeps={model.pca_sweeps}", flush=True)
base = outs[(False, False)]
for k, v in outs.items():
    print(k, "max |d| vs serial cold:", float((v - base).abs().max()), "equal" if torch.equal(v, base) else "")
