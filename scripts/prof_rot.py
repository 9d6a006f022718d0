import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, stream_ptr
lib = _lib.lib()
dev = torch.device("cuda", 0)
st = stream_ptr(dev)
c = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for k in (1, 16):
    ws = torch.empty(lib.optex_rotations_workspace_bytes(c, k), dtype=torch.uint8, device=dev)
    rots = torch.empty(k, c, c, device=dev)
    for _ in range(2):
        call("optex_random_rotations", ptr(rots), c, k, 1, 0, None, ptr(ws), ws.numel(), st)
    torch.cuda.synchronize()
