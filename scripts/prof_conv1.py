import sys; sys.path.insert(0, ".")
import torch, optimaltextures_b200 as ob
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.relu(torch.randn(1, 1024, 1024, 64, device="cuda", generator=g)); s = torch.relu(torch.randn(1, 1024, 1024, 64, device="cuda", generator=g))
rots = ob.random_rotations(64, 2, "cuda", seed=1)
for i in range(2): p = ob.optimal_transport(p, s, "cdf", rotation=rots[i])
torch.cuda.synchronize()
