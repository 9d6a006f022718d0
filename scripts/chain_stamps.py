"""Clock stamps of the cooperative covariance-chain kernel (cov_chain.cu) over one pca OT step: cycles between the grid
barriers of CTA 0.  Usage: python scripts/chain_stamps.py [hw] [c] [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import optimaltextures_b200 as ob
from optimaltextures_b200._runtime import call, ptr

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 32
c = int(sys.argv[2]) if len(sys.argv) > 2 else 320
mode = sys.argv[3] if len(sys.argv) > 3 else "pca"
g = torch.Generator(device="cuda").manual_seed(0)
decay = (0.985 ** torch.arange(c, device="cuda")) * 30.0
p = torch.randn(1, hw, hw, c, device="cuda", generator=g) * decay * 0.7 + 0.3
s = torch.randn(1, hw, hw, c, device="cuda", generator=g) * decay + 0.5
ob.ot_loop(p, s, mode, 3)
torch.cuda.synchronize()
stamps = torch.zeros(256, dtype=torch.int64, device="cuda")
call("optex_debug_chain_stamps", ptr(stamps), 256)
ob.ot_loop(p, s, mode, 2)            # the second iteration (style side reused) overwrites the first one's stamps
torch.cuda.synchronize()
call("optex_debug_chain_stamps", None, 0)
t = stamps.cpu()
t = t[t > 0]
cut = next((i for i in range(1, len(t)) if t[i] < t[i - 1]), len(t))   # stale stamps of the longer first iteration
t = t[:cut]
d = (t[1:] - t[:-1]).tolist()
print(f"{mode} hw={hw} c={c}: {len(t)} stamps, total {int(t[-1] - t[0])} clk; intervals: {d}")
