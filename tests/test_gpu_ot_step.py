"""GPU parity: rotation GEMMs, the whole OT step, the inner loop and the rotation generator."""
import numpy as np
import pytest
import torch

from optimaltextures_b200 import _lib
from oracle import ot_oracle, rotation as rot_oracle, sort_oracle

pytestmark = pytest.mark.gpu

T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
OT_CASES = ["sq16", "ragged23", "batch2", "batch2_s1", "wide64", "rgb_ties"]
GEMM_MODES = ["fp32", "auto"]   # auto = tcgen05 3xTF32 where the shape allows, fp32 SIMT otherwise

# fp32 tolerance of one OT step on unit-scale features (stated here, used below):
#   rotation GEMMs: |err| <= 2e-5 * scale  (fp32 accumulation-order differences only)
#   per-channel modes are DISCONTINUOUS in their input at bin edges / rank swaps, so a last-bit
#   difference in a rotated value can move one output by a whole bin: the bulk must agree to
#   BULK_TOL and at most OUTLIER_FRAC of the elements may differ by more.
GEMM_TOL = 2e-5
BULK_TOL = 2e-4
OUTLIER_FRAC = 5e-3
COV_TOL = 5e-4


@pytest.fixture(scope="module")
def ob():
    import optimaltextures_b200 as ob

    return ob


@pytest.fixture(params=GEMM_MODES)
def gemm_mode(request, ob):
    ob.set_gemm_mode(request.param)
    yield request.param
    ob.set_gemm_mode("auto")


def close_in_bulk(out, ref, tol=BULK_TOL, frac=OUTLIER_FRAC):
    out, ref = np.asarray(out, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = max(1.0, float(np.abs(ref).max()))
    bad = np.abs(out - ref) > tol * scale
    assert bad.mean() <= frac, f"{bad.mean():.4%} of elements differ by more than {tol * scale:g}"


@pytest.mark.parametrize("n,c", [(64, 16), (120, 23), (4096, 64), (16384, 512), (1000, 181), (333, 3),
                                 (65536, 64), (262144, 64), (65536, 128), (16384, 320), (36864, 512)])
def test_rotation_gemms_match_fp32_matmul(ob, gemm_mode, n, c):
    g = torch.Generator().manual_seed(n + c)
    x = torch.relu(torch.randn(n, c, generator=g))
    r = T(rot_oracle.haar_rotation_qr(c, 1)).float()
    ref = (x.double() @ r.double())
    xt = ob.rotate_forward(x.cuda(), r.cuda())
    assert xt.shape == (c, n)
    scale = float(ref.abs().max())
    assert float((xt.cpu().double().T - ref).abs().max()) <= GEMM_TOL * scale
    back = ob.rotate_inverse(xt, r.cuda())
    assert float((back.cpu().double() - ref @ r.double().T).abs().max()) <= 2 * GEMM_TOL * scale
    # fused content blend (optex.py:117)
    content = torch.randn(n, c, generator=g)
    blended = ob.rotate_inverse(xt, r.cuda(), content=content.cuda(), content_strength=0.25)
    expect = back.cpu() + 0.25 * (content - back.cpu())
    np.testing.assert_array_equal(blended.cpu().numpy(), expect.numpy())


@pytest.mark.parametrize("n,c", [(128, 32), (4096, 64), (16384, 512), (4096, 320)])
def test_tf32_single_pass_mode(ob, n, c):
    """OPTEX_GEMM_TF32 = the reference's own CUDA default (allow_tf32, optex.py:248-249): 10-bit mantissas."""
    g = torch.Generator().manual_seed(n + c)
    x = torch.relu(torch.randn(n, c, generator=g))
    r = T(rot_oracle.haar_rotation_qr(c, 1)).float()
    ref = x.double() @ r.double()
    ob.set_gemm_mode("tf32")
    try:
        xt = ob.rotate_forward(x.cuda(), r.cuda())
    finally:
        ob.set_gemm_mode("auto")
    err = float((xt.cpu().double().T - ref).abs().max()) / float(ref.abs().max())
    assert 1e-5 < err < 3e-3      # really TF32 (not fp32), and no worse than TF32


def test_forced_tensor_core_mode_rejects_unsupported_shapes(ob):
    ob.set_gemm_mode("tf32x3")
    try:
        with pytest.raises(ValueError):
            ob.rotate_forward(torch.zeros(64, 23).cuda(), torch.eye(23).cuda())     # C % 32 != 0
    finally:
        ob.set_gemm_mode("auto")


@pytest.mark.parametrize("name", OT_CASES)
def test_ot_step_cdf_vs_reference_golden(ob, golden, gemm_mode, name):
    g = golden("ot_step")
    t, s, rot = T(g[f"{name}_t"]), T(g[f"{name}_s"]), T(g[f"{name}_rot"])
    out = ob.optimal_transport(t.cuda(), s.cuda(), "cdf", rotation=rot.cuda())
    assert out.shape == t.shape and out.is_contiguous()
    close_in_bulk(out.cpu().numpy(), g[f"{name}_out_cdf"])


@pytest.mark.parametrize("name", OT_CASES)
def test_ot_step_cdf_is_bit_exact_given_the_same_rotated_features(ob, golden, name):
    """Decomposed parity: GPU rotate -> (GPU cdf vs CPU oracle cdf on the SAME rotated values)."""
    g = golden("ot_step")
    t, s, rot = T(g[f"{name}_t"]), T(g[f"{name}_s"]), T(g[f"{name}_rot"]).float()
    rp, rs = ob.rotate_forward(t.cuda(), rot.cuda()), ob.rotate_forward(s.cuda(), rot.cuda())
    got = ob.cdf_match(rp, rs).cpu()
    ref = ot_oracle.cdf_match_channels(rp.cpu(), rs.cpu())
    np.testing.assert_array_equal(got.numpy(), ref.numpy())


@pytest.mark.parametrize("name", OT_CASES)
def test_ot_step_sort_vs_oracle(ob, golden, gemm_mode, name):
    g = golden("ot_step")
    t, s, rot = T(g[f"{name}_t"]), T(g[f"{name}_s"]), T(g[f"{name}_rot"])
    out = ob.optimal_transport(t.cuda(), s.cuda(), "sort", rotation=rot.cuda())
    close_in_bulk(out.cpu().numpy(), sort_oracle.ot_step_sort(t, s, rot).numpy())


@pytest.mark.parametrize("mode", ["cdf", "sort"])
def test_inner_loop_with_content(ob, golden, mode):
    g = golden("inner_loop")
    t, s, content, rots = T(g["t"]), T(g["s"]), T(g["content"]), T(g["rots"])
    strength = ot_oracle.content_strength_for_layer(0.2, 1)
    got = ob.ot_loop(t.cuda(), s.cuda(), mode, 3, rotations=rots.float().cuda(), content=content.cuda(),
                     content_strength=strength)
    # step-by-step through optimal_transport gives the identical result (same kernels, same order)
    p = t.cuda()
    for r in rots:
        p = ob.optimal_transport(p, s.cuda(), mode, rotation=r.cuda(), content=content.cuda(),
                                 content_strength=strength)
    assert torch.equal(got, p)
    if mode == "cdf":
        close_in_bulk(got.cpu().numpy(), g["out_cdf"], tol=5e-4, frac=2e-2)


@pytest.mark.parametrize("mode", ["cdf", "sort"])
@pytest.mark.parametrize("device_rotations", [False, True])
def test_ot_loop_fused_rotations(ob, mode, device_rotations):
    """ot_loop without content fuses `un-rotate with R_i, rotate with R_{i+1}` into one rotation by R_i^T R_{i+1}
    (two GEMMs per iteration instead of three).  A chain of matcher steps amplifies last-bit differences (the cdf
    map is discontinuous at bin edges and the next rotation spreads a flipped bin over all channels; SURVEY H4:
    parity is a per-step notion), so the criterion is self-calibrating: the fused loop must be no further from the
    step-by-step loop than the step-by-step loop is from ITSELF under a different fp32-grade GEMM arithmetic
    (fp32 FFMA vs 3xTF32)."""
    c = 64
    g = torch.Generator().manual_seed(3)
    p = torch.relu(torch.randn(1, 64, 64, c, generator=g)).cuda()
    s = torch.relu(1.3 * torch.randn(1, 48, 80, c, generator=g) + 0.2).cuda()

    def stepwise(rots):
        ref = p
        for r in rots:
            ref = ob.optimal_transport(ref, s, mode, rotation=r)
        return ref

    for iters in (2, 5):
        rots = ob.random_rotations(c, iters, "cuda", seed=21, first_counter=100)
        if device_rotations:
            got = ob.ot_loop(p, s, mode, iters, seed=21, first_counter=100)
        else:
            got = ob.ot_loop(p, s, mode, iters, rotations=rots)
        ref = stepwise(rots)
        ob.set_gemm_mode("fp32")
        try:
            ref_fp32 = stepwise(rots)
        finally:
            ob.set_gemm_mode("auto")
        dist = lambda x, y: float((x - y).abs().mean())
        assert dist(got, ref) <= 1.5 * dist(ref, ref_fp32) + 1e-6, (iters, dist(got, ref), dist(ref, ref_fp32))
        # and it is a real transport: the result's marginals along the last rotation are the style's
        q = torch.tensor([0.1, 0.5, 0.9], device="cuda")
        a = torch.quantile((got.reshape(-1, c) @ rots[-1])[:, :8], q, dim=0)
        b = torch.quantile((s.reshape(-1, c) @ rots[-1])[:, :8], q, dim=0)
        assert float((a - b).abs().max()) < (0.05 if mode == "cdf" else 5e-3)


def test_ot_step_host_buffers(ob, golden):
    g = golden("ot_step")
    t, s, rot = T(g["wide64_t"]), T(g["wide64_s"]), T(g["wide64_rot"]).float()
    dev_out = ob.optimal_transport(t.cuda(), s.cuda(), "cdf", rotation=rot.cuda())
    host_out = ob.optimal_transport_host(t, s, rot, "cdf")
    assert not host_out.is_cuda
    assert torch.equal(host_out, dev_out.cpu())


def test_ot_step_host_async_double_buffered(ob, golden):
    """Two pipelined slots on two streams give the same results as the synchronous call."""
    g = golden("ot_step")
    t, s, rot = T(g["wide64_t"]).pin_memory(), T(g["wide64_s"]).pin_memory(), T(g["wide64_rot"]).float().pin_memory()
    ref = {m: ob.optimal_transport_host(t, s, rot, m) for m in ("cdf", "chol")}
    streams = [torch.cuda.Stream() for _ in range(2)]
    outs = [torch.empty_like(t).pin_memory() for _ in range(2)]
    for rep in range(3):
        for i, m in enumerate(("cdf", "chol")):
            ob.optimal_transport_host(t, s, rot, m, out=outs[i], slot=i, stream=streams[i])
        torch.cuda.synchronize()
        for i, m in enumerate(("cdf", "chol")):
            assert torch.equal(outs[i], ref[m])


def test_headline_shape_properties(ob):
    """conv4_1 @ 1024^2 (N = 16384, C = 512): properties that hold at any size."""
    gen = torch.Generator(device="cuda").manual_seed(0)
    p = torch.relu(torch.randn(1, 128, 128, 512, device="cuda", generator=gen))
    s = torch.relu(1.3 * torch.randn(1, 128, 128, 512, device="cuda", generator=gen) + 0.2)
    r = ob.random_rotation(512, "cuda", seed=1, counter=0)
    for mode in ("cdf", "sort"):
        out = ob.optimal_transport(p, s, mode, rotation=r)
        assert torch.isfinite(out).all()
        # sliced-OT property: along every rotated direction the output's marginal is the style's
        ro, rs = (out.reshape(-1, 512) @ r), (s.reshape(-1, 512) @ r)
        q = torch.tensor([0.1, 0.5, 0.9], device="cuda")
        err = (torch.quantile(ro[:, :8], q, dim=0) - torch.quantile(rs[:, :8], q, dim=0)).abs().max()
        assert float(err) < (0.05 if mode == "cdf" else 2e-3)
    # identity rotation + matching a block to itself with `sort` is the identity map
    ob.set_gemm_mode("fp32")          # exact products with the identity (3xTF32 drops bits beyond 2^-22)
    try:
        same = ob.optimal_transport(p, p, "sort", rotation=torch.eye(512, device="cuda"))
    finally:
        ob.set_gemm_mode("auto")
    assert torch.equal(same, p)


# ------------------------------------------------------------------ rotations
@pytest.fixture(params=["fp64", "fp32"])
def rot_precision(ob, request):
    """Both arithmetic types of the Householder construction (optex_set_rotation_precision)."""
    prev = ob.set_rotation_precision(request.param)
    yield request.param
    ob.set_rotation_precision(prev)


@pytest.mark.parametrize("n", [1, 2, 3, 23, 64, 181, 512, 700])
def test_rotation_matches_householder_oracle(ob, rot_precision, n):
    rng = np.random.RandomState(n)
    gauss = rng.normal(size=(max(n - 1, 0), n))
    got = ob.random_rotation(n, "cuda", gauss=T(gauss)).cpu().double().numpy()
    ref = rot_oracle.haar_rotation_householder(gauss) if n > 1 else np.ones((1, 1))
    # fp64 construction rounds once (fp32 storage); the fp32 one (the reference's impl="torch" arithmetic,
    # optex.py:150-164) carries the rounding of up to n - 1 successive reflections
    np.testing.assert_allclose(got, ref, atol=1e-6)


@pytest.mark.parametrize("n", [3, 64, 512, 1024])
def test_rotation_is_special_orthogonal_and_reproducible(ob, rot_precision, n):
    a = ob.random_rotation(n, "cuda", seed=7, counter=3)
    b = ob.random_rotation(n, "cuda", seed=7, counter=3)
    c = ob.random_rotation(n, "cuda", seed=7, counter=4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    ad = a.double()
    # fp64 construction: the only error is the fp32 rounding of the entries; fp32 construction: n - 1 rounded
    # reflections (measured 1.3e-6 at n = 512)
    orth_tol = 5e-7 if rot_precision == "fp64" else 5e-6
    assert float((ad @ ad.T - torch.eye(n, device="cuda", dtype=torch.float64)).abs().max()) < orth_tol
    assert abs(float(torch.linalg.det(ad)) - 1.0) < (1e-5 if rot_precision == "fp64" else 1e-4)


def test_rotation_batch_equals_single_draws(ob, rot_precision):
    """The batched draw (different rows-per-warp configuration) gives the same matrices as one draw at a time."""
    batch = ob.random_rotations(256, 24, "cuda", seed=5, first_counter=10)
    for i in (0, 7, 23):
        one = ob.random_rotation(256, "cuda", seed=5, counter=10 + i)
        assert float((batch[i] - one).abs().max()) < (2e-7 if rot_precision == "fp64" else 2e-6)


def test_rotation_haar_statistics(ob):
    """Distributional parity with scipy's special_ortho_group (optex.py:149): for Haar SO(n) every
    entry has mean 0 and variance 1/n, and E[trace] = 0."""
    n, reps = 16, 400
    mats = torch.stack([ob.random_rotation(n, "cuda", seed=11, counter=i) for i in range(reps)]).double()
    assert abs(float(mats.mean())) < 3.0 / np.sqrt(reps * n * n * n)
    assert abs(float(mats.var()) * n - 1.0) < 0.02
    tr = mats.diagonal(dim1=1, dim2=2).sum(1)
    assert abs(float(tr.mean())) < 4.0 / np.sqrt(reps)
    assert abs(float(tr.var()) - 1.0) < 0.3        # Var[trace] = 1 for Haar O(n)/SO(n)


def test_implicit_rotation_stream_is_pooled_and_reproducible(ob):
    """optimal_transport() without rotation= draws like the reference (optex.py:168); the draws come from a pool
    filled 16 at a time, and the stream is a pure function of manual_seed() and the call order."""
    g = torch.Generator(device="cpu").manual_seed(1)
    p = torch.relu(torch.randn(1, 16, 16, 64, generator=g)).cuda()
    s = torch.relu(torch.randn(1, 16, 16, 64, generator=g)).cuda()
    ob.manual_seed(11)
    a = [ob.optimal_transport(p, s, "cdf") for _ in range(18)]        # crosses a pool refill
    ob.manual_seed(11)
    b = [ob.optimal_transport(p, s, "cdf") for _ in range(18)]
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert not torch.equal(a[0], a[1])
    rots = ob.random_rotations(64, 16, "cuda", seed=11, first_counter=0)
    for i in (0, 5, 15):
        assert torch.equal(a[i], ob.optimal_transport(p, s, "cdf", rotation=rots[i]))
    rots2 = ob.random_rotations(64, 2, "cuda", seed=11, first_counter=16)
    assert torch.equal(a[17], ob.optimal_transport(p, s, "cdf", rotation=rots2[1]))


def test_prepared_rotation_is_bit_identical(ob):
    """optex_rotation_prepare only moves the hi / lo split of R out of the GEMM calls: same bits, fewer launches."""
    g = torch.Generator(device="cpu").manual_seed(2)
    x = torch.randn(8192, 256, generator=g).cuda()
    r = ob.random_rotation(256, "cuda", seed=3, counter=0)
    lib = _lib.lib()
    plain_f = ob.rotate_forward(x, r)
    plain_i = ob.rotate_inverse(plain_f, r)
    n0 = lib.optex_launch_count()
    ob.rotate_forward(x, r); ob.rotate_inverse(plain_f, r)
    unprepared = lib.optex_launch_count() - n0
    with ob.prepared_rotation(r) as rr:
        n0 = lib.optex_launch_count()
        prep_f = ob.rotate_forward(x, rr)
        prep_i = ob.rotate_inverse(prep_f, rr)
        prepared = lib.optex_launch_count() - n0
    assert torch.equal(plain_f, prep_f) and torch.equal(plain_i, prep_i)
    assert prepared == unprepared - 2          # the two per-call split kernels are gone
    after = ob.rotate_forward(x, r)            # released: back to the self-contained call
    assert torch.equal(after, plain_f)
