"""GEMM determinism / accuracy matrix through optex_pca_project (B K-major with transpose=True, MN-major without):
several tiles per CTA, K beyond 512, each N tile width (OPTEX_FORCE_BN is read once per process).  Test infrastructure."""
import os
import sys

import torch

sys.path.insert(0, ".")
import optimaltextures_b200 as ob

print("FORCE_BN", os.environ.get("OPTEX_FORCE_BN"), "NO_A_TMEM", os.environ.get("OPTEX_NO_A_TMEM"), flush=True)
g = torch.Generator().manual_seed(0)
for (n, kdim, cdim, tr) in [(16384, 576, 128, True), (16384, 512, 128, True), (16384, 576, 64, True),
                            (65536, 64, 64, True), (65536, 576, 64, True), (16384, 576, 128, False),
                            (32768, 512, 512, True), (32768, 512, 512, False), (16384, 320, 320, False)]:
    x = torch.randn(n, kdim, generator=g).cuda()
    if tr:   # out[n, c] = x[n, k] V[c, k]^T
        v = torch.randn(cdim, kdim, generator=g).cuda()
        ref = (x.double() @ v.double().T)
    else:    # out[n, k2] = x[n, c] V[c, k2]
        v = torch.randn(kdim, cdim, generator=g).cuda()
        ref = (x.double() @ v.double())
    outs = [ob.pca_project(x, v, transpose=tr) for _ in range(4)]
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    errs = [float((o.double() - ref).abs().max()) / scale for o in outs]
    bad_rows = ((outs[0].double() - ref).abs() / scale > 1e-4).any(1).nonzero().flatten()
    tiles = sorted(set((bad_rows // 128).tolist()))
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print(f"M={n} K={kdim} N={cdim} B={'K' if tr else 'MN'}-major: err {max(errs):.1e} same {same} bad m-tiles {tiles[:12]} "
          f"({len(tiles)})", flush=True)
