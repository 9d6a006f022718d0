"""A-via-TMEM 3xTF32 GEMM: accuracy against fp64 and hot timing (forward transposed, inverse from channel-major)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
from optimaltextures_b200._runtime import call, ptr, stream_ptr
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
ob.set_gemm_mode(mode)
for (n, c) in [(16384, 512), (16384 + 96, 256), (65536, 128), (8192, 512)]:
    x = torch.randn(n, c, device=dev) * 3 + 1
    r = ob.random_rotation(c, device=dev)
    out = torch.full((c, n), 7.0, device=dev)
    call("optex_rotate_forward", ptr(x), ptr(r), ptr(out), n, c, stream_ptr(dev))
    torch.cuda.synchronize()
    ref = (x.double() @ r.double()).T
    print(f"fwd n={n} c={c}: err={float((out - ref).abs().max()):.3e} sevens={int((out == 7).sum())}", flush=True)
    o2 = torch.full((n, c), 7.0, device=dev)
    call("optex_rotate_inverse", ptr(out), ptr(r), ptr(o2), n, c, None, 0.0, stream_ptr(dev))
    torch.cuda.synchronize()
    ref2 = out.double().T @ r.double().T
    print(f"inv n={n} c={c}: err={float((o2 - ref2).abs().max()):.3e} roundtrip={float((o2 - x).abs().max()):.3e}", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("fwd", lambda: call("optex_rotate_forward", ptr(x), ptr(r), ptr(out), n, c, stream_ptr(dev))),
                     ("inv", lambda: call("optex_rotate_inverse", ptr(out), ptr(r), ptr(o2), n, c, None, 0.0, stream_ptr(dev)))):
        for _ in range(5):
            fn()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"  {name} hot {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)

# cold inputs: cycle over 8 distinct A operands (8 x 32 MB > L2) at the headline shape
n, c = 16384, 512
xs = [torch.randn(n, c, device=dev) for _ in range(8)]
outs = [torch.empty(c, n, device=dev) for _ in range(8)]
o2s = [torch.empty(n, c, device=dev) for _ in range(2)]
r = ob.random_rotation(c, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda i: call("optex_rotate_forward", ptr(xs[i % 8]), ptr(r), ptr(outs[i % 8]), n, c, stream_ptr(dev))),
                 ("inv", lambda i: call("optex_rotate_inverse", ptr(outs[i % 8]), ptr(r), ptr(o2s[i % 2]), n, c, None, 0.0, stream_ptr(dev)))):
    for i in range(8):
        fn(i)
    e0.record()
    for i in range(48):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    print(f"  {name} cold (8 operand sets) {e0.elapsed_time(e1) / 48 * 1e3:.1f} us", flush=True)
