"""GPU regression: the tensor-core GEMM when one CTA walks SEVERAL tiles, for every N-tile width the dispatcher picks,
K below / above 512 and both B majors - through `optex_pca_project` (optex.py:110, :120), the thinnest C-ABI wrapper
around it.  Found by scripts/debug_gemm_multi.py: the TMEM-A operand path with 64-wide tiles corrupted isolated
32-row groups of second tiles, non-deterministically; these shapes pin the dispatcher's choice for them.

Tolerance: 3xTF32 products with fp32 accumulation vs an fp64 matmul: 2e-5 * max|ref| (K <= 1152); and every repeat
must be bit-identical."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # n, k, c, transpose (transpose: out[n, c] = x[n, k] V[c, k]^T, B K-major; else out[n, c] = x[n, k] V[k, c])
    (16384, 576, 128, True), (16384, 512, 128, True), (16384, 576, 64, True), (65536, 64, 64, True),
    (65536, 576, 64, True), (16384, 576, 128, False), (32768, 512, 512, True), (32768, 512, 512, False),
    (16384, 320, 320, False), (65536, 1152, 128, True), (20000, 96, 64, False),
]


@pytest.mark.parametrize("n,k,c,transpose", SHAPES)
def test_many_tiles_per_cta(n, k, c, transpose):
    import optimaltextures_b200 as ob

    g = torch.Generator().manual_seed(n + k + c)
    x = torch.randn(n, k, generator=g).cuda()
    if transpose:
        v = torch.randn(c, k, generator=g).cuda()
        ref = x.double() @ v.double().T
    else:
        v = torch.randn(k, c, generator=g).cuda()
        ref = x.double() @ v.double()
    outs = [ob.pca_project(x, v, transpose=transpose) for _ in range(4)]
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    for o in outs:
        assert float((o.double() - ref).abs().max()) <= 2e-5 * scale
    assert all(torch.equal(outs[0], o) for o in outs[1:])
