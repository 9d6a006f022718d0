"""OT-step throughput at the five VGG layer shapes of a 1024^2 image (SURVEY.md 8d sweep), all hist modes."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob

shapes = [("conv5_1", 64, 512), ("conv4_1", 128, 512), ("conv3_1", 256, 256), ("conv2_1", 512, 128), ("conv1_1", 1024, 64)]
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cdf", "sort", "chol", "pca", "sym"]
res = {}
for name, hw, c in shapes:
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.relu(torch.randn(1, hw, hw, c, device="cuda", generator=g))
    s = torch.relu(1.3 * torch.randn(1, hw, hw, c, device="cuda", generator=g) + 0.2)
    rots = ob.random_rotations(c, 8, "cuda", seed=1)
    for mode in modes:
        try:
            for i in range(2):
                ob.optimal_transport(p, s, mode, rotation=rots[i])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 8
            e0.record()
            for i in range(iters):
                out = ob.optimal_transport(p, s, mode, rotation=rots[i])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            res[f"{name}:{mode}"] = round(1e3 / ms, 1)
            ok = bool(torch.isfinite(out).all())
            print(f"{name} N={hw*hw:8d} C={c:4d} {mode:5s} {ms*1e3:9.1f} us/step {1e3/ms:9.1f} it/s finite={ok}", flush=True)
        except Exception as ex:
            res[f"{name}:{mode}"] = f"{type(ex).__name__}: {str(ex)[:80]}"
            print(f"{name} {mode}: {type(ex).__name__}: {str(ex)[:120]}", flush=True)
print(json.dumps(res))
