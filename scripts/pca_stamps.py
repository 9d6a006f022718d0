"""Where a global round of the blocked-order Jacobi solver (pca.cu) spends its time: clock64 stamps of CTA 1
(optex_debug_pca_stamps) over the first rounds of a cold solve.  Usage: python scripts/pca_stamps.py [c] [n]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, stream_ptr, workspace

c = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda")
lib = _lib.lib()
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(n, c, generator=g) @ (torch.randn(c, c, generator=g) * (2.0 / c ** 0.5)) + 0.3).to(dev)
vec, sig = torch.empty(c, c, device=dev), torch.empty(c, device=dev)
kd = torch.zeros(2, dtype=torch.int32, device=dev)
ws = workspace(dev, lib.optex_fit_pca_workspace_bytes(n, c))
NR = 400
stamps = torch.zeros(NR, 5, dtype=torch.int64, device=dev)


def run():
    call("optex_fit_pca_warm", ptr(x), n, c, ptr(vec), ptr(sig), ptr(kd), None, 0, C.c_void_p(kd.data_ptr() + 4),
         ptr(ws), ws.numel(), stream_ptr(dev))


run()
torch.cuda.synchronize()
call("optex_debug_pca_stamps", ptr(stamps), NR)
run()
torch.cuda.synchronize()
call("optex_debug_pca_stamps", None, 0)
s = stamps.cpu()
rounds_per_sweep = (c + 15) // 16 * 2 - 1
ok = [i for i in range(NR - 1) if i % rounds_per_sweep != 0 and s[i, 4] > 0 and s[i + 1, 0] > 0]
d = lambda a, b: float(torch.tensor([int(s[i, b] - s[i, a]) for i in ok], dtype=torch.float64).median())
print(f"c={c} n={n} sweeps={int(kd[1])} rounds/sweep={rounds_per_sweep}; median clocks over {len(ok)} cross-pair rounds:")
print(f"  load rows + norms      {d(0, 1):8.0f}")
print(f"  8 sub-rounds           {d(1, 2):8.0f}")
print(f"  store rows             {d(2, 3):8.0f}")
print(f"  grid barrier           {d(3, 4):8.0f}")
print(f"  whole round            {d(0, 4):8.0f}")
