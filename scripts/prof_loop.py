import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
mode = sys.argv[1] if len(sys.argv) > 1 else "cdf"
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.relu(torch.randn(1, 128, 128, 512, device="cuda", generator=g))
s = torch.relu(1.3 * torch.randn(1, 128, 128, 512, device="cuda", generator=g) + 0.2)
out = ob.ot_loop(p, s, mode, 3, seed=1, first_counter=0)
torch.cuda.synchronize()
print("done", float(out.sum()))
