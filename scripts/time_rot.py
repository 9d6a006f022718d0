"""Device time of the batched rotation draw (optex_random_rotations) against the batch size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, stream_ptr
lib = _lib.lib()
dev = torch.device("cuda", 0)
st = stream_ptr(dev)
for c in (512, 256, 128, 64):
    for k in (1, 4, 8, 16, 30, 64):
        ws = torch.empty(lib.optex_rotations_workspace_bytes(c, k), dtype=torch.uint8, device=dev)
        rots = torch.empty(k, c, c, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            call("optex_random_rotations", ptr(rots), c, k, 1, 0, None, ptr(ws), ws.numel(), st)
        e0.record()
        for _ in range(5):
            call("optex_random_rotations", ptr(rots), c, k, 1, 0, None, ptr(ws), ws.numel(), st)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 * 1e3
        print(f"c={c} batch={k}: {t:.0f} us total, {t / k:.1f} us per rotation", flush=True)
