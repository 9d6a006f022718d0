"""A few OT steps at the headline shape, for ncu (launch list / --set full captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
mode = sys.argv[1] if len(sys.argv) > 1 else "cdf"
gemm = sys.argv[2] if len(sys.argv) > 2 else "auto"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ob.set_gemm_mode(gemm)
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.relu(torch.randn(1, 128, 128, 512, device="cuda", generator=g))
s = torch.relu(1.3 * torch.randn(1, 128, 128, 512, device="cuda", generator=g) + 0.2)
rots = ob.random_rotations(512, steps, "cuda", seed=1)
for i in range(steps):
    p = ob.optimal_transport(p, s, mode, rotation=rots[i])
torch.cuda.synchronize()
print("done", float(p.sum()))
