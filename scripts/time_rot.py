"""Device time of the batched rotation draw (optex_random_rotations) against batch size and arithmetic."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, stream_ptr
lib = _lib.lib()
dev = torch.device("cuda", 0)
st = stream_ptr(dev)
for prec in ("fp64", "fp32"):
    ob.set_rotation_precision(prec)
    for c in (512, 256, 64):
        for k in (1, 8, 16, 30, 64):
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
            rd = rots[0].double()
            orth = float((rd @ rd.T - torch.eye(c, device=dev, dtype=torch.float64)).abs().max())
            print(f"{prec} c={c} batch={k}: {t:.0f} us total, {t / k:.1f} us per rotation, |RR^T - I| = {orth:.1e}", flush=True)
