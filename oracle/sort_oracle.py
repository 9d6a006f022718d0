"""`sort` mode oracle - exact 1-D optimal transport per rotated channel.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This mode is NOT in the reference (its per-channel
mode is the 256-bin `cdf`, histmatch.py:49-69) - parity versus the reference is UNPINNED / not
applicable.  It is the north-star's "rotate, sort, match, unrotate" step and
this file is its definition:

    idx  = stable argsort of the target channel          (ties broken by pixel index; -0.0 == +0.0)
    ss   = ascending sort of the source channel
    rank r of target element idx[r] receives  ss[q(r)],  q(r) = ((2r + 1) * m) // (2n)

q is the mid-point quantile map in integer arithmetic (q(r) == r when m == n), so permutation
indices and outputs are bit-exact quantities.
"""
from __future__ import annotations

import torch
from torch import Tensor


def quantile_index(n: int, m: int) -> Tensor:
    r = torch.arange(n, dtype=torch.int64)
    return ((2 * r + 1) * m) // (2 * n)


def sort_match_channels(target: Tensor, source: Tensor):
    """target [c, n], source [c, m] -> (matched [c, n], idx [c, n] int64)."""
    n, m = target.shape[1], source.shape[1]
    idx = torch.sort(target, dim=1, stable=True).indices
    ss = torch.sort(source, dim=1).values
    out = torch.empty_like(target)
    out.scatter_(1, idx, ss[:, quantile_index(n, m)])
    return out, idx


def ot_step_sort(pastiche: Tensor, style: Tensor, rotation: Tensor) -> Tensor:
    """optex.py:167-177 with `hist_match` replaced by the exact sort matcher."""
    rotation = rotation.to(pastiche)
    c = pastiche.shape[-1]
    rp = (pastiche @ rotation).reshape(-1, c).T.contiguous()
    rs = (style @ rotation).reshape(-1, c).T.contiguous()
    m, _ = sort_match_channels(rp, rs)
    return (m.T @ rotation.T).reshape(pastiche.shape)
