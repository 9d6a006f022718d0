import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
from optimaltextures_b200._runtime import call, ptr, stream_ptr
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for (n, c) in [(128, 32), (256, 64), (128, 256)]:
    for mode in ("tf32",):
        ob.set_gemm_mode(mode)
        x = torch.randn(n, c, device=dev)
        r = torch.eye(c, device=dev)
        out = torch.full((c, n), 7.0, device=dev)
        call("optex_rotate_forward", ptr(x), ptr(r), ptr(out), n, c, stream_ptr(dev))
        torch.cuda.synchronize()
        print(f"fwd n={n} c={c}: out sum={float(out.sum()):.3f} sevens={int((out==7).sum())} zeros={int((out==0).sum())} "
              f"err_vs_xT={float((out - x.T).abs().max()):.3e}")
        print(" out[0,:8]", out[0, :8].tolist()); print(" x[:8,0] ", x[:8, 0].tolist())
        r2 = torch.randn(c, c, device=dev)
        call("optex_rotate_forward", ptr(x), ptr(r2), ptr(out), n, c, stream_ptr(dev))
        torch.cuda.synchronize()
        ref = (x @ r2).T
        print(f"  randR err={float((out-ref).abs().max()):.3e} ref max={float(ref.abs().max()):.3f}")
        mt = torch.randn(c, n, device=dev)
        o2 = torch.full((n, c), 7.0, device=dev)
        call("optex_rotate_inverse", ptr(mt), ptr(r), ptr(o2), n, c, None, 0.0, stream_ptr(dev))
        torch.cuda.synchronize()
        print(f"inv n={n} c={c}: sevens={int((o2==7).sum())} zeros={int((o2==0).sum())} err_vs_mtT={float((o2 - mt.T).abs().max()):.3e}")
        call("optex_rotate_inverse", ptr(mt), ptr(r2), ptr(o2), n, c, None, 0.0, stream_ptr(dev))
        torch.cuda.synchronize()
        ref = mt.T @ r2.T
        print(f"  randR err={float((o2-ref).abs().max()):.3e} ref max={float(ref.abs().max()):.3f}")
