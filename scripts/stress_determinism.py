"""Run the product's synthesis loop repeatedly (same injected rotations / noise) and report every run that differs
from the first - a race detector for the whole pipeline (tests/test_gpu_texture.py::test_end_to_end_sanity asserts
two equal runs).  Between runs unrelated work perturbs allocator / L2 / workspace state."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import optimaltextures_b200 as ob
from optimaltextures_b200 import texture
from oracle import texture_cases

import test_gpu_texture as T

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for name in ("mix_content_chol_opt", "synth_pca"):
    first = None
    bad = 0
    for i in range(reps):
        model, out = T._run_product(ob, name)
        if first is None:
            first = out.clone()
        elif not torch.equal(out, first):
            bad += 1
            print(f"{name}: run {i} differs: max |d| = {float((out - first).abs().max()):.3g}", flush=True)
        # perturb: unrelated steps of other shapes / modes, fresh allocations
        g = torch.Generator().manual_seed(i)
        p = torch.relu(torch.randn(1, 40 + 8 * (i % 5), 64, 128, generator=g)).cuda()
        for mode in ("cdf", "chol", "pca"):
            ob.optimal_transport(p, p.flip(1), mode)
        junk = torch.randn(1 << (20 + i % 4), device="cuda")
        del junk, p
    print(f"{name}: {bad} of {reps - 1} repeat runs differ from the first", flush=True)
