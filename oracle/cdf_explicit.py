"""numpy restatement of `cdf_match` / `interp` with the fp32 arithmetic written out.

TEST INFRASTRUCTURE (see oracle/__init__.py).  No torch in here: this file states,
operation by operation and rounding by rounding, what the CUDA kernels in
`optimaltextures_b200/csrc/cdf_match.cu` must reproduce BIT-EXACTLY.

Reference: /root/reference/histmatch.py:49-92.  The arithmetic lives in torch
(un-vendored, un-pinned by the reference; pinned here against torch 2.11.0 CPU,
AVX-512 build, by `tests/test_oracle_golden.py` + `tests/test_cdf_explicit.py`):

* `torch.histc(x, bins, lo, hi)`        (histmatch.py:57-58)
      bin = int64( (x - lo) * float(bins) / (hi - lo) ),  bin == bins -> bins-1,
      every operation rounded to fp32; if lo == hi the range becomes [lo-1, hi+1].
* `torch.linspace(lo, hi, bins+1)[1:]`  (histmatch.py:59)
      step = (hi - lo) / float(bins);
      edge[i] (i = 1..bins) = fma(step, i, lo)            if i <  (bins+1)//2
                            = fma(-step, bins - i, hi)    otherwise   (so edge[bins] == hi)
* `cumsum` of fp32 counts is exact below 2**24; `/ cdf[-1]` is one fp32 division.
* `torch.searchsorted(xp, x)` = first index i with xp[i] >= x  (left).
* interp: separate fp32 sub, sub, div, sub, mul, add - torch never fuses them.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def fma32(a, b, c):
    """Correctly-rounded fp32 fused multiply-add of fp32 arrays (a*b + c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    p = a * b  # exact: 24b x 24b fits in 53b
    s = p + c  # rounded to double
    # TwoSum error term (exact): err = (p + c) - s
    bb = s - p
    err = (p - (s - bb)) + (c - bb)
    # Double rounding can only go wrong when s is EXACTLY an fp32 rounding midpoint while the
    # true value (s + err) is not: then the tie must be broken by sign(err), not to-even.
    r = s.astype(F)
    d = s - r.astype(np.float64)
    toward = np.where(d > 0, np.inf, -np.inf).astype(F)
    r2 = np.nextafter(r, toward)
    is_mid = (d != 0) & (2.0 * d == (r2.astype(np.float64) - r.astype(np.float64)))
    fix = is_mid & (err != 0)
    if np.any(fix):
        s = np.where(fix, np.nextafter(s, np.where(err > 0, np.inf, -np.inf)), s)
        r = s.astype(F)
    return r


def histc_bin_index(x, lo, hi, bins: int):
    """Bin index rule of torch.histc on CPU (HistogramKernel.cpp, linear-bin path)."""
    x = np.asarray(x, dtype=F)
    lo = F(lo)
    hi = F(hi)
    if lo == hi:
        lo = F(lo - F(1))
        hi = F(hi + F(1))
    with np.errstate(all="ignore"):
        q = ((x - lo).astype(F) * F(bins)).astype(F) / F(hi - lo)
        b = q.astype(np.int64)
    b = np.where(b == bins, bins - 1, b)
    return np.clip(b, 0, bins - 1)  # clip: memory safety for NaN/degenerate input only


def histc(x, lo, hi, bins: int):
    return np.bincount(histc_bin_index(x, lo, hi, bins), minlength=bins).astype(F)


def linspace_upper_edges(lo, hi, bins: int):
    """torch.linspace(lo, hi, bins+1)[1:] on CPU (RangeFactoriesKernel.cpp), fp32."""
    lo = F(lo)
    hi = F(hi)
    step = F(F(hi - lo) / F(bins))
    i = np.arange(1, bins + 1)
    half = (bins + 1) // 2
    first = fma32(np.full(bins, step, F), i.astype(F), np.full(bins, lo, F))
    second = fma32(np.full(bins, -step, F), (bins - i).astype(F), np.full(bins, hi, F))
    return np.where(i < half, first, second).astype(F)


def interp_backward(x, xp, fp):
    """histmatch.py:72-92, in explicit fp32 steps."""
    x = np.asarray(x, dtype=F)
    xp = np.asarray(xp, dtype=F)
    fp = np.asarray(fp, dtype=F)
    last = xp.shape[0] - 1
    i = np.minimum(np.searchsorted(xp, x, side="left"), last)  # == searchsorted (never > last, §3.5)
    j = np.minimum(i + 1, last)
    with np.errstate(all="ignore"):
        slope = ((fp[j] - fp[i]).astype(F) / (xp[j] - xp[i]).astype(F)).astype(F)
        f = ((slope * (x - xp[i]).astype(F)).astype(F) + fp[i]).astype(F)
        bad = ~np.isfinite(f)
        f2 = ((slope * (x - xp[j]).astype(F)).astype(F) + fp[j]).astype(F)
    f = np.where(bad, f2, f)
    bad2 = ~np.isfinite(f)
    f = np.where(bad2, fp[i], f)
    return f.astype(F)


def cdf_tables(t, s, bins: int = 256):
    """lo, hi, upper edges and the remap table for one channel (histmatch.py:52-67)."""
    t = np.asarray(t, dtype=F)
    s = np.asarray(s, dtype=F)
    lo = F(min(t.min(), s.min()))
    hi = F(max(t.max(), s.max()))
    th = histc(t, lo, hi, bins)
    sh = histc(s, lo, hi, bins)
    edges = linspace_upper_edges(lo, hi, bins)
    tc = np.cumsum(th, dtype=np.float64).astype(F)  # exact integers < 2**24
    sc = np.cumsum(sh, dtype=np.float64).astype(F)
    tc = (tc / tc[-1]).astype(F)
    sc = (sc / sc[-1]).astype(F)
    remap = interp_backward(tc, sc, edges)
    return lo, hi, edges, remap, th, sh


def cdf_match(target, source, bins: int = 256):
    """histmatch.py:49-69; target [c, n], source [c, m] -> [c, n] (fp32)."""
    target = np.asarray(target, dtype=F)
    source = np.asarray(source, dtype=F)
    out = np.empty_like(target)
    for ch in range(target.shape[0]):
        _, _, edges, remap, _, _ = cdf_tables(target[ch], source[ch], bins)
        out[ch] = interp_backward(target[ch], edges, remap)
    return out
