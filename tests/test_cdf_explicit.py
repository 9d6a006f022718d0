"""The explicit-arithmetic numpy restatement vs the torch ops the reference calls (CPU, this build)."""
import numpy as np
import pytest
import torch

from oracle import cdf_explicit, ot_oracle


def test_histc_rule_matches_torch():
    rng = np.random.RandomState(0)
    for bins in (256, 100, 7):
        for _ in range(40):
            x = (rng.randn(4000) * rng.uniform(0.1, 5)).astype(np.float32)
            lo, hi = x.min(), x.max()
            ref = torch.histc(torch.from_numpy(x), bins, float(lo), float(hi)).numpy()
            np.testing.assert_array_equal(cdf_explicit.histc(x, lo, hi, bins), ref)


def test_linspace_rule_matches_torch():
    rng = np.random.RandomState(1)
    for bins in (256, 100, 7):
        for _ in range(200):
            lo = np.float32(rng.randn() * 3 - 2)
            hi = np.float32(lo + abs(rng.randn()) * 10 + 1e-3)
            ref = torch.linspace(float(lo), float(hi), bins + 1)[1:].numpy()
            got = cdf_explicit.linspace_upper_edges(lo, hi, bins)
            np.testing.assert_array_equal(got, ref)
            assert got[-1] == hi


@pytest.mark.parametrize("seed", range(4))
def test_cdf_match_matches_torch_oracle(seed):
    g = torch.Generator().manual_seed(seed)
    c, n, m = 5, 3000 + 517 * seed, 2500
    t = torch.randn(c, n, generator=g) * 1.7
    s = torch.relu(torch.randn(c, m, generator=g) * 1.3 + 0.2)   # many ties at 0 -> flat CDF segments
    ref = ot_oracle.cdf_match_channels(t, s).numpy()
    np.testing.assert_array_equal(cdf_explicit.cdf_match(t.numpy(), s.numpy()), ref)


def test_fma32_is_correctly_rounded():
    from fractions import Fraction
    rng = np.random.RandomState(3)
    a = rng.randn(2000).astype(np.float32)
    b = rng.randint(1, 256, 2000).astype(np.float32)
    c = (rng.randn(2000) * 4).astype(np.float32)
    got = cdf_explicit.fma32(a, b, c)
    for i in range(2000):
        exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        lo_, hi_ = np.nextafter(got[i], np.float32(-np.inf)), np.nextafter(got[i], np.float32(np.inf))
        err = abs(Fraction(float(got[i])) - exact)
        assert err <= abs(Fraction(float(lo_)) - exact) and err <= abs(Fraction(float(hi_)) - exact)
