"""Why is the first process on a fresh box slower?  Repeats the bench's timed region and prints device time and CPU enqueue time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, stream_ptr, workspace
lib = _lib.lib()
dev = torch.device("cuda", 0)
n, c, K = 16384, 512, 50
sets = [(torch.relu(torch.randn(1, 128, 128, c, device=dev)), torch.relu(torch.randn(1, 128, 128, c, device=dev))) for _ in range(4)]
outs = [torch.empty_like(sets[0][0]) for _ in range(2)]
rots = ob.random_rotations(c, K, "cuda", seed=1)
ws = workspace(dev, lib.optex_ot_workspace_bytes(n, n, c, 3)); st = stream_ptr(dev)
def step(i):
    p, s = sets[i % 4]
    call("optex_ot_step", ptr(p), ptr(s), ptr(rots[i % K]), ptr(outs[i % 2]), 1, n, 1, n, c, 3, 1.0, None, 0.0, ptr(ws), ws.numel(), st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"rep {rep}: device {e0.elapsed_time(e1) / K * 1e3:.1f} us/step, cpu enqueue {(t1 - t0) / K * 1e6:.1f} us/step", flush=True)
    if rep == 3:
        time.sleep(2.0)
