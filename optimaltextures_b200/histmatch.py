"""Drop-in for the reference's ``histmatch`` module (/root/reference/histmatch.py), CUDA only.

Same names, argument meaning and return shapes as the reference:

    hist_match(target, source, mode="chol", eps=1)   histmatch.py:5-46
    cdf_match(target, source, bins=256)              histmatch.py:49-69
    interp(x, xp, fp)                                histmatch.py:72-92

plus ``sort_match`` (the north-star's exact per-channel 1-D OT, not in the reference).
Every function enqueues hand-written sm_100a kernels from liboptex_b200.so on the current
torch CUDA stream; nothing synchronises and nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _lib
from ._runtime import call, f32c, ptr, require_cuda, stream_ptr, workspace


def hist_match(target: Tensor, source: Tensor, mode: str = "chol", eps: float = 1) -> Tensor:
    """reference: histmatch.py:5-46.  target [b,h,w,c], source [bs,hs,ws,c] (NHWC) -> [b,h,w,c].

    Returns a fresh contiguous NHWC tensor (the reference returns a permuted view; its callers only
    matmul / blend it, optex.py:175,203)."""
    dev = require_cuda(target, source)
    if target.dim() != 4 or source.dim() != 4 or target.shape[-1] != source.shape[-1]:
        raise ValueError(f"hist_match expects NHWC tensors with equal channels, got {tuple(target.shape)} "
                         f"and {tuple(source.shape)}")
    t, s = f32c(target), f32c(source)
    b, h, w, c = t.shape
    bs, hs, ws_, _ = s.shape
    out = torch.empty_like(t)
    if t.numel() == 0:
        return out
    m = _lib.mode_id(mode)
    nbytes = _lib.lib().optex_hist_match_workspace_bytes(b * h * w, bs * hs * ws_, c, m)
    wsb = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        call("optex_hist_match", ptr(t), ptr(s), ptr(out), b, h * w, bs, hs * ws_, c, m, float(eps), ptr(wsb),
             wsb.numel(), stream_ptr(dev))
    return out.to(target.dtype)


def cdf_match(target: Tensor, source: Tensor, bins: int = 256, return_tables: bool = False):
    """reference: histmatch.py:49-69.  target [c, n], source [c, m] (channel-major) -> [c, n].

    Bit-exact with the reference's torch-CPU arithmetic.  ``return_tables`` additionally returns the
    per-channel (upper bin edges, remapped CDF) as a [c, 2, bins] tensor."""
    dev = require_cuda(target, source)
    if target.dim() != 2 or source.dim() != 2 or target.shape[0] != source.shape[0]:
        raise ValueError(f"cdf_match expects [c, n] and [c, m], got {tuple(target.shape)} and {tuple(source.shape)}")
    t, s = f32c(target), f32c(source)
    c, n = t.shape
    out = torch.empty_like(t)
    tables = torch.empty(c, 2, bins, dtype=torch.float32, device=dev) if return_tables else None
    nbytes = _lib.lib().optex_cdf_match_workspace_bytes(c, bins)
    wsb = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        call("optex_cdf_match", ptr(t), ptr(s), ptr(out), c, n, s.shape[1], int(bins), ptr(tables), ptr(wsb),
             wsb.numel(), stream_ptr(dev))
    return (out, tables) if return_tables else out


def sort_match(target: Tensor, source: Tensor, return_perm: bool = False):
    """Exact 1-D OT per channel (oracle/sort_oracle.py): target [c, n], source [c, m] -> [c, n].

    ``return_perm`` additionally returns the stable argsort of every target channel (int32 [c, n])."""
    dev = require_cuda(target, source)
    if target.dim() != 2 or source.dim() != 2 or target.shape[0] != source.shape[0]:
        raise ValueError(f"sort_match expects [c, n] and [c, m], got {tuple(target.shape)} and {tuple(source.shape)}")
    t, s = f32c(target), f32c(source)
    c, n = t.shape
    out = torch.empty_like(t)
    perm = torch.empty(c, n, dtype=torch.int32, device=dev) if return_perm else None
    nbytes = _lib.lib().optex_sort_match_workspace_bytes(c, n, s.shape[1])
    wsb = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        call("optex_sort_match", ptr(t), ptr(s), ptr(out), c, n, s.shape[1], ptr(perm), ptr(wsb), wsb.numel(),
             stream_ptr(dev))
    return (out, perm) if return_perm else out


def interp(x: Tensor, xp: Tensor, fp: Tensor) -> Tensor:
    """reference: histmatch.py:72-92 (NOT np.interp: backward extrapolation + two non-finite fallbacks)."""
    dev = require_cuda(x, xp, fp)
    if xp.dim() != 1 or fp.shape != xp.shape or xp.numel() < 1:
        raise ValueError("interp expects 1-D xp and fp of equal, non-zero length")
    xc, xpc, fpc = f32c(x), f32c(xp), f32c(fp)
    out = torch.empty_like(xc)
    with torch.cuda.device(dev):
        call("optex_interp", ptr(xc), ptr(xpc), ptr(fpc), ptr(out), xc.numel(), xpc.numel(), stream_ptr(dev))
    return out
