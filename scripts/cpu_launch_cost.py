import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, stream_ptr, workspace
lib = _lib.lib()
dev = torch.device("cuda", 0)
n, c = 16384, 512
p = torch.relu(torch.randn(1, 128, 128, c, device=dev)); s = torch.relu(torch.randn(1, 128, 128, c, device=dev))
r = ob.random_rotation(c, "cuda", seed=1, counter=0); out = torch.empty_like(p)
ws = workspace(dev, lib.optex_ot_workspace_bytes(n, n, c, 3)); st = stream_ptr(dev)
def step():
    call("optex_ot_step", ptr(p), ptr(s), ptr(r), ptr(out), 1, n, 1, n, c, 3, 1.0, None, 0.0, ptr(ws), ws.numel(), st)
for _ in range(5): step()
torch.cuda.synchronize()
K = 40
t0 = time.perf_counter()
for _ in range(K): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e6*(t1-t0)/K:.1f} us/step (CPU), total {1e6*(t2-t0)/K:.1f} us/step")
xt = torch.empty(c, n, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(K): call("optex_rotate_forward", ptr(p), ptr(r), ptr(xt), n, c, st)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"rotate_forward: enqueue {1e6*(t1-t0)/K:.1f} us/call, total {1e6*(t2-t0)/K:.1f}")
