"""Per-CTA duration of the rotation GEMMs (debug stamps of optex_debug_gemm_trace: clock64 + globaltimer at entry and
exit of every CTA): cycles per CTA, effective SM clock during the kernel, spread over the grid; and event-timed
duration of the forward / inverse rotation back to back.  OPTEX_CTA_GROUP=1 selects the single-CTA tiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import optimaltextures_b200 as ob
from optimaltextures_b200 import _lib

lib = _lib.lib()
lib.optex_set_pdl(0)
n, c = int(os.environ.get("GN", 16384)), int(os.environ.get("GC", 512))
g = torch.Generator().manual_seed(0)
xs = [torch.relu(torch.randn(n, c, generator=g)).cuda() for _ in range(6)]      # 200 MB > L2: A comes from HBM
r = ob.random_rotation(c, "cuda", seed=1, counter=0)
ref = (xs[0].double() @ r.double()).T
print("OPTEX_CTA_GROUP", os.environ.get("OPTEX_CTA_GROUP"), flush=True)
with ob.prepared_rotation(r) as rr:
    out = ob.rotate_forward(xs[0], rr)
    print("forward max err / scale:", float((out.double() - ref).abs().max() / ref.abs().max()), flush=True)
    back = ob.rotate_inverse(out, rr)
    print("inverse max err / scale:", float((back.double() - xs[0].double()).abs().max() / xs[0].abs().max()), flush=True)
    outs = [ob.rotate_forward(x, rr) for x in xs]
    torch.cuda.synchronize()
    assert all(torch.equal(ob.rotate_forward(xs[i], rr), outs[i]) for i in range(6)), "not deterministic"
    for name, fn in (("forward", lambda i: ob.rotate_forward(xs[i % 6], rr)), ("inverse", lambda i: ob.rotate_inverse(outs[i % 6], rr))):
        for _ in range(6):
            fn(_)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(60):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 60 * 1e3:.1f} us per launch (60 back to back, inputs cycle over 200 MB)", flush=True)
        tr = torch.zeros(1024 + 4 * 32 * 4, dtype=torch.int64, device="cuda")
        tr[95] = 0x7ace
        lib.optex_debug_gemm_trace(tr.data_ptr())
        fn(1)
        torch.cuda.synchronize()
        lib.optex_debug_gemm_trace(None)
        full = tr.cpu()
        t = full[128:128 + 4 * 148].view(148, 4)
        t = t[t[:, 0] != 0]
        cyc = (t[:, 2] - t[:, 0]).double()
        ns = (t[:, 3] - t[:, 1]).double()
        span = float(t[:, 3].max() - t[:, 1].min())
        print(f"  {len(t)} CTAs: cycles min/median/max {cyc.min():.0f}/{cyc.median():.0f}/{cyc.max():.0f}; "
              f"ns min/median/max {ns.min():.0f}/{ns.median():.0f}/{ns.max():.0f}; grid span {span:.0f} ns; "
              f"clock ~ {float((cyc / ns).median()) * 1e3:.0f} MHz", flush=True)
        rs = full[1024:].view(4, 32, 4)
        base = int(rs[rs != 0].min())
        print("  CTA 0, per k block (cycles): producer[wait_afree wait_empty] | mma[t_start wait_B wait_A issue] | "
              "converter[t_start wait_raw wait_ta_empty work]")
        for kb in range(32):
            pr, mm, cv = rs[0, kb], rs[1, kb], rs[2, kb]
            if int(mm[0]) == 0:
                continue
            print(f"   kb{kb:2d}: prod [{int(pr[1] - pr[0]):5d} {int(pr[3] - pr[2]):5d}] | mma [{int(mm[0]) - base:6d} "
                  f"{int(mm[1] - mm[0]):5d} {int(mm[2] - mm[1]):5d} {int(mm[3] - mm[2]):4d}] | conv [{int(cv[0]) - base:6d} "
                  f"{int(cv[1] - cv[0]):5d} {int(cv[2] - cv[1]):5d} {int(cv[3] - cv[2]):5d}]")
        for t_ in range(2):
            e = rs[3, t_]
            if int(e[0]):
                print(f"   epilogue warp 8, tile {t_}: wait from {int(e[0]) - base}, accumulator ready {int(e[1]) - base}, "
                      f"drained {int(e[2]) - base}, stores done {int(e[3]) - base}")
            for ch in range(2):
                e = rs[3, 4 + t_ * 2 + ch]
                if int(e[0]):
                    print(f"      chunk {ch}: start {int(e[0]) - base}, buffer free {int(e[1]) - base}, staged {int(e[2]) - base}, "
                          f"store issued {int(e[3]) - base}")
