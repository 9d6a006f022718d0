"""GPU parity: per-channel matchers and interp through the C-ABI vs the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import cdf_explicit, ot_oracle, sort_oracle

pytestmark = pytest.mark.gpu

T = lambda a: torch.from_numpy(np.ascontiguousarray(a))


@pytest.fixture(scope="module")
def ob():
    import optimaltextures_b200 as ob

    return ob


def dev(a):
    return (a if isinstance(a, torch.Tensor) else T(a)).cuda()


# ------------------------------------------------------------------ interp
def test_interp_known_answers(ob, golden):
    g = golden("kat_interp_cdf")
    for k in (0, 1):
        out = ob.interp(dev(g[f"interp{k}_x"]), dev(g[f"interp{k}_xp"]), dev(g[f"interp{k}_fp"]))
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"interp{k}_out"])


def test_interp_random_bit_exact(ob):
    g = torch.Generator().manual_seed(3)
    xp = torch.sort(torch.rand(300, generator=g)).values
    xp[40:44] = xp[40]                      # ties -> 0/0 slopes -> both fallbacks
    xp[-1] = 1.0
    fp = torch.randn(300, generator=g)
    x = torch.rand(5000, generator=g).clamp(max=1.0)
    x[:300] = xp                            # exact hits
    ref = ot_oracle.interp_backward(x, xp, fp)
    np.testing.assert_array_equal(ob.interp(dev(x), dev(xp), dev(fp)).cpu().numpy(), ref.numpy())


# ------------------------------------------------------------------ cdf_match
@pytest.mark.parametrize("k", [0, 1, 2])
def test_cdf_match_golden(ob, golden, k):
    g = golden("kat_interp_cdf")
    out = ob.cdf_match(dev(g[f"cdf{k}_t"]), dev(g[f"cdf{k}_s"]))
    np.testing.assert_array_equal(out.cpu().numpy(), g[f"cdf{k}_out"])


CDF_SHAPES = [
    (5, 3000, 2500),      # ragged, scalar (unaligned) path
    (3, 1, 1),            # single element
    (7, 4096, 4096),      # vector path
    (2, 70001, 333),      # > one private-histogram flush per thread, odd sizes
    (64, 16384, 9216),    # conv4_1-like block, many channels
    (1, 300000, 262144),  # multi-CTA split of one channel
]


@pytest.mark.parametrize("c,n,m", CDF_SHAPES)
def test_cdf_match_bit_exact_vs_oracle(ob, c, n, m):
    g = torch.Generator().manual_seed(c * 7919 + n)
    t = torch.randn(c, n, generator=g) * 1.7
    s = torch.relu(torch.randn(c, m, generator=g) * 1.3 + 0.2)       # ties at 0 -> flat CDF segments
    out, tables = ob.cdf_match(dev(t), dev(s), return_tables=True)
    out, tables = out.cpu().numpy(), tables.cpu().numpy()
    nchk = c if n * c <= 2_000_000 else 2
    for ch in range(nchk):
        _, _, edges, remap, _, _ = cdf_explicit.cdf_tables(t[ch].numpy(), s[ch].numpy())
        np.testing.assert_array_equal(tables[ch, 0], edges)
        np.testing.assert_array_equal(tables[ch, 1], remap)
    ref = ot_oracle.cdf_match_channels(t[:nchk], s[:nchk]).numpy()
    np.testing.assert_array_equal(out[:nchk], ref)


def test_cdf_match_edge_cases(ob):
    # constant channel (hi == lo), quantised data with massive ties, negative zero, huge/small scales
    t = torch.stack([torch.full((100,), 3.25), torch.round(torch.rand(100) * 4) / 4, torch.zeros(100),
                     torch.randn(100) * 1e20, torch.randn(100) * 1e-20])
    t[2, ::2] = -0.0
    s = torch.stack([torch.full((60,), 3.25), torch.round(torch.rand(60) * 4) / 4, torch.zeros(60),
                     torch.randn(60) * 1e20, torch.randn(60) * 1e-20])
    ref = ot_oracle.cdf_match_channels(t, s).numpy()
    out = ob.cdf_match(dev(t), dev(s)).cpu().numpy()
    np.testing.assert_array_equal(out, ref)
    # empty target: nothing to do; empty source: error like the reference (min of empty)
    assert ob.cdf_match(torch.empty(3, 0).cuda(), torch.rand(3, 5).cuda()).shape == (3, 0)
    with pytest.raises(ValueError):
        ob.cdf_match(torch.rand(3, 5).cuda(), torch.empty(3, 0).cuda())


@pytest.mark.parametrize("bins", [7, 100, 512, 1024])
def test_cdf_match_other_bin_counts(ob, bins):
    g = torch.Generator().manual_seed(5)
    t, s = torch.randn(3, 5000, generator=g), torch.randn(3, 4000, generator=g) * 2 + 1
    ref = ot_oracle.cdf_match_channels(t, s, bins).numpy()
    np.testing.assert_array_equal(ob.cdf_match(dev(t), dev(s), bins=bins).cpu().numpy(), ref)


def test_cdf_match_full_size_properties(ob):
    """BASELINE size (conv4_1 @ 1024^2: 512 x 16384): size-independent properties."""
    g = torch.Generator(device="cuda").manual_seed(0)
    t = torch.randn(512, 16384, device="cuda", generator=g)
    s = torch.relu(torch.randn(512, 16384, device="cuda", generator=g) * 1.3 + 0.2)
    out = ob.cdf_match(t, s)
    assert torch.isfinite(out).all()
    # the map is a function of the value within a channel: equal inputs -> equal outputs, and
    # matching a channel to itself leaves the upper bin edges fixed points
    order = torch.argsort(t, dim=1)
    ts, os_ = torch.gather(t, 1, order), torch.gather(out, 1, order)
    same = ts[:, 1:] == ts[:, :-1]
    assert (os_[:, 1:][same] == os_[:, :-1][same]).all()
    # spot-check 3 channels against the CPU oracle, bit for bit
    for ch in (0, 255, 511):
        ref = ot_oracle.cdf_match_channels(t[ch:ch + 1].cpu(), s[ch:ch + 1].cpu())
        assert torch.equal(out[ch:ch + 1].cpu(), ref)


# ------------------------------------------------------------------ sort_match
SORT_SHAPES = [(4, 512, 512), (3, 1000, 777), (5, 16384, 16384), (2, 4096, 16000), (6, 1, 1), (3, 37, 5000),
               (2, 9000, 100)]


@pytest.mark.parametrize("c,n,m", SORT_SHAPES)
def test_sort_match_bit_exact(ob, c, n, m):
    g = torch.Generator().manual_seed(n * 31 + m)
    t = torch.randn(c, n, generator=g)
    s = torch.relu(torch.randn(c, m, generator=g) * 1.3 + 0.2)
    ref, idx = sort_oracle.sort_match_channels(t, s)
    out, perm = ob.sort_match(dev(t), dev(s), return_perm=True)
    np.testing.assert_array_equal(perm.cpu().numpy().astype(np.int64), idx.numpy())
    np.testing.assert_array_equal(out.cpu().numpy(), ref.numpy())


def test_sort_match_ties_are_stable(ob):
    g = torch.Generator().manual_seed(9)
    t = torch.round(torch.rand(3, 5000, generator=g) * 255) / 255       # 8-bit colour data (H8)
    t[0, ::3] = -0.0
    t[0, 1::3] = 0.0
    s = torch.round(torch.rand(3, 4000, generator=g) * 255) / 255
    ref, idx = sort_oracle.sort_match_channels(t, s)
    out, perm = ob.sort_match(dev(t), dev(s), return_perm=True)
    np.testing.assert_array_equal(perm.cpu().numpy().astype(np.int64), idx.numpy())
    np.testing.assert_array_equal(out.cpu().numpy(), ref.numpy())


def test_sort_match_full_size_properties(ob):
    g = torch.Generator(device="cuda").manual_seed(1)
    t = torch.randn(512, 16384, device="cuda", generator=g)
    s = torch.randn(512, 16384, device="cuda", generator=g) * 2 + 1
    out, perm = ob.sort_match(t, s, return_perm=True)
    # perm is a permutation that sorts t; out is a permutation of s (n == m) in t's order
    assert torch.equal(torch.sort(perm.long(), dim=1).values, torch.arange(16384, device="cuda").expand(512, -1))
    assert (torch.diff(torch.gather(t, 1, perm.long()), dim=1) >= 0).all()
    assert torch.equal(torch.sort(out, dim=1).values, torch.sort(s, dim=1).values)
    assert torch.equal(torch.gather(out, 1, perm.long()), torch.sort(s, dim=1).values)


LARGE_SORT_SHAPES = [(3, 16385, 16384), (2, 40000, 70001), (2, 65536, 65536), (1, 300000, 1000), (2, 1000, 100000),
                     (1, 1048576, 1048576)]


@pytest.mark.parametrize("c,n,m", LARGE_SORT_SHAPES)
def test_sort_match_large_channels_bit_exact(ob, c, n, m):
    """Channels longer than the on-chip capacity (16384): chunk sort + stable merge passes."""
    g = torch.Generator().manual_seed(n + m)
    t = torch.randn(c, n, generator=g)
    t[:, ::7] = torch.round(t[:, ::7] * 4) / 4          # ties across chunk boundaries
    s = torch.relu(torch.randn(c, m, generator=g) * 1.3 + 0.2)
    ref, idx = sort_oracle.sort_match_channels(t, s)
    out, perm = ob.sort_match(dev(t), dev(s), return_perm=True)
    np.testing.assert_array_equal(perm.cpu().numpy().astype(np.int64), idx.numpy())
    np.testing.assert_array_equal(out.cpu().numpy(), ref.numpy())
