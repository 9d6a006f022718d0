import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimaltextures_b200 as ob
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.relu(torch.randn(1, 128, 128, 512, device="cuda", generator=g))
s = torch.relu(1.3 * torch.randn(1, 128, 128, 512, device="cuda", generator=g) + 0.2)
for mode in sys.argv[1].split(","):
    ob.ot_loop(p, s, mode, 4, seed=1, first_counter=0); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence = torch.zeros(1, device="cuda")
        e0.record(); out = ob.ot_loop(p, s, mode, 32, seed=1, first_counter=0); fence.add_(1); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 32)
    print(f"ot_loop {mode}: {best*1000:.1f} us/iter {1/best*1000:.0f} it/s  fusion={'off' if os.environ.get('OPTEX_NO_LOOP_FUSION') else 'on'}")
