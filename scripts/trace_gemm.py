import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
from optimaltextures_b200 import _lib
lib = _lib.lib()
n, c = 16384, 512
x = torch.relu(torch.randn(n, c, device="cuda")); r = ob.random_rotation(c, "cuda", seed=1, counter=0)
for mode in ("tf32x3", "tf32"):
    ob.set_gemm_mode(mode)
    for _ in range(3): ob.rotate_forward(x, r)
    tr = torch.zeros(96, dtype=torch.int64, device="cuda")
    lib.optex_debug_gemm_trace(tr.data_ptr())
    ob.rotate_forward(x, r); torch.cuda.synchronize()
    lib.optex_debug_gemm_trace(None)
    full = tr.cpu(); t = full[:64].view(16, 4); t0 = int(t[0, 0])
    names = {64: "entry", 65: "after pdl_wait", 66: "setup done", 67: "exit", 68: "epi0 start", 69: "epi0 end", 70: "epi1 start", 71: "epi1 end", 76: "tile1 mma kb0", 77: "tile1 mma last kb"}
    print({v: int(full[k]) - t0 for k, v in names.items() if int(full[k])})
    print(mode, "kb: tma_issue raw_landed split_done mma_start (cycles from first issue)")
    for kb in range(16):
        print(kb, [int(v) - t0 if int(v) else None for v in t[kb]])
