"""torch-CPU oracle of the reference hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function restates one reference function and cites it.  The third-party
arithmetic (torch 2.11 CPU: matmul, linalg.cholesky/eigh, inverse, histc,
linspace, cumsum, searchsorted) is *called*, not re-derived, so on this torch
build the results are bit-identical to the reference; `tests/test_oracle_golden.py`
pins that against vectors produced by the real reference.

The one deliberate difference: the rotation is an ARGUMENT.  The reference draws
it inside `optimal_transport` from numpy's global RNG (optex.py:149,168), which
makes value parity impossible; injecting it is the only way to compare.
"""
from __future__ import annotations

import torch
from torch import Tensor

COV_MODES = ("chol", "pca", "sym")


# --------------------------------------------------------------------------- interp
def interp_backward(x: Tensor, xp: Tensor, fp: Tensor) -> Tensor:
    """reference: histmatch.py:72-92 (`interp`).

    NOT np.interp: with i = searchsorted_left(xp, x), j = min(i+1, len-1) it
    evaluates the segment (i, j) *to the right* of x, anchored at i; non-finite
    results are re-anchored at j, and whatever is still non-finite becomes fp[i].
    """
    last = xp.numel() - 1
    i = torch.searchsorted(xp, x)
    j = (i + 1).clamp(0, last)
    slope = (fp[j] - fp[i]) / (xp[j] - xp[i])
    out = slope * (x - xp[i]) + fp[i]
    bad = ~torch.isfinite(out)
    if bad.any():
        jb = j[bad]
        out[bad] = slope[bad] * (x[bad] - xp[jb]) + fp[jb]
        bad2 = ~torch.isfinite(out)
        if bad2.any():
            out[bad2] = fp[i[bad2]]
    return out


# --------------------------------------------------------------------------- cdf
def cdf_tables(t: Tensor, s: Tensor, bins: int = 256):
    """Per-channel intermediates of histmatch.py:52-67 (lo, hi, edges, remap)."""
    lo = torch.min(t.min(), s.min())
    hi = torch.max(t.max(), s.max())
    t_hist = torch.histc(t, bins, lo, hi)
    s_hist = torch.histc(s, bins, lo, hi)
    edges = torch.linspace(lo, hi, bins + 1)[1:]
    t_cdf = t_hist.cumsum(0)
    t_cdf = t_cdf / t_cdf[-1]
    s_cdf = s_hist.cumsum(0)
    s_cdf = s_cdf / s_cdf[-1]
    remap = interp_backward(t_cdf, s_cdf, edges)
    return lo, hi, edges, remap, t_hist, s_hist


def cdf_match_channels(target: Tensor, source: Tensor, bins: int = 256) -> Tensor:
    """reference: histmatch.py:49-69 (`cdf_match`); target [c, n], source [c, m]."""
    out = torch.empty_like(target)
    tc = target.contiguous()
    for ch in range(tc.shape[0]):
        t, s = tc[ch], source[ch]
        _, _, edges, remap, _, _ = cdf_tables(t, s, bins)
        out[ch] = interp_backward(t, edges, remap)
    return out


# --------------------------------------------------------------------------- hist_match
def _centered_cov(x_cbhw: Tensor, eps: float):
    """histmatch.py:16-18 / :20-22: per-(c,b) mean over (h,w); pooled covariance + eps*I."""
    c = x_cbhw.shape[0]
    mu = x_cbhw.mean((2, 3), keepdim=True)
    flat = (x_cbhw - mu).view(c, -1)
    cov = flat @ flat.T / flat.shape[1] + eps * torch.eye(c, device=x_cbhw.device)
    return mu, flat, cov


def _psd_sqrt(mat: Tensor) -> Tensor:
    """V sqrt(diag(w)) V^T with eigh(UPLO='U'), as at histmatch.py:30-31."""
    w, v = torch.linalg.eigh(mat, UPLO="U")
    return v @ torch.sqrt(torch.diag(w)) @ v.T


def hist_match_nhwc(target: Tensor, source: Tensor, mode: str = "chol", eps: float = 1) -> Tensor:
    """reference: histmatch.py:5-46 (`hist_match`); target [b,h,w,c], source [bs,hs,ws,c]."""
    t = target.permute(3, 0, 1, 2)
    s = source.permute(3, 0, 1, 2)
    c, b, h, w = t.shape
    if mode == "cdf":
        m = cdf_match_channels(t.view(c, -1), s.view(c, -1)).view(c, b, h, w)
        return m.permute(1, 2, 3, 0)

    _, t_flat, cov_t = _centered_cov(t, eps)
    mu_s, _, cov_s = _centered_cov(s, eps)
    if mode == "chol":
        l_t = torch.linalg.cholesky(cov_t)
        l_s = torch.linalg.cholesky(cov_s)
        m = l_s @ torch.inverse(l_t) @ t_flat
    elif mode == "pca":
        q_t = _psd_sqrt(cov_t)
        q_s = _psd_sqrt(cov_s)
        m = q_s @ torch.inverse(q_t) @ t_flat
    else:  # "sym" and, like the reference (histmatch.py:36), any unknown string
        q_t = _psd_sqrt(cov_t)
        mid = _psd_sqrt(q_t @ cov_s @ q_t)
        m = torch.inverse(q_t) @ mid @ torch.inverse(q_t) @ t_flat
    m = m.view(c, b, h, w) + mu_s
    return m.permute(1, 2, 3, 0)


# --------------------------------------------------------------------------- OT step
def ot_step(pastiche: Tensor, style: Tensor, rotation: Tensor, mode: str) -> Tensor:
    """reference: optex.py:167-177 (`optimal_transport`) with the rotation injected."""
    rotation = rotation.to(pastiche)
    rp = pastiche @ rotation
    rs = style @ rotation
    return hist_match_nhwc(rp, rs, mode=mode) @ rotation.T


def content_strength_for_layer(content_strength: float, l: int) -> float:
    """optex.py:116: strength = content_strength / 2**(4-l), applied for l <= 2 only."""
    return content_strength / 2 ** (4 - l)


def content_blend_(pastiche: Tensor, content: Tensor, strength: float) -> Tensor:
    """optex.py:117: in-place  p += strength * (content - p)."""
    pastiche += strength * (content - pastiche)
    return pastiche


def ot_loop(pastiche, style, rotations, mode, content=None, content_strength=0.0, l=0):
    """reference: optex.py:112-117, the inner loop for one layer (rotations injected)."""
    for rot in rotations:
        pastiche = ot_step(pastiche, style, rot, mode)
        if content is not None and l <= 2:
            content_blend_(pastiche, content, content_strength_for_layer(content_strength, l))
    return pastiche


# --------------------------------------------------------------------------- fit_pca (next row, §8f-2)
def fit_pca(tensor: Tensor):
    """reference: optex.py:180-190 (`fit_pca`)."""
    a = tensor.reshape(-1, tensor.shape[-1]) - tensor.mean()
    _, sv, v = torch.svd(a)
    k = (torch.cumsum(sv / torch.sum(sv), dim=0) > 0.9).max(0).indices.squeeze()
    v = v[:, :k]
    return tensor @ v, v
