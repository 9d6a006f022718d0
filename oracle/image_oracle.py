"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py) for the image-space glue of one synthesis pass.

  resize              util.py:105-106  - torch's own CPU `interpolate` is the reference arithmetic (called, not
                                         re-derived); `resize_explicit` writes the same separable antialiased
                                         bicubic filter out in numpy fp32 - it is the arithmetic csrc/image.cu
                                         implements and tests/test_image_oracle.py pins it to torch here.
  rgb_to_hls / hls_to_rgb / lightness_transfer
                      optex.py:126-128 - kornia.color.hls is NOT vendored by the reference and absent from this
                                         image: restated from kornia's published formulas, PARITY UNPINNED.
  mix_layer           optex.py:197-204 - one layer of mix_style_features given the two hist_match results.
  recentre            optex.py:76
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


def resize(x: Tensor, size) -> Tensor:
    """reference: util.py:105-106."""
    return F.interpolate(x, size=tuple(size), mode="bicubic", align_corners=False, antialias=True)


def _cubic_aa(x: np.ndarray) -> np.ndarray:
    a = np.float32(-0.5)
    x = np.abs(x).astype(np.float32)
    near = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + np.float32(1)
    far = (((x - np.float32(5)) * x + np.float32(8)) * x - np.float32(4)) * a
    return np.where(x < 1, near, np.where(x < 2, far, np.float32(0))).astype(np.float32)


def resize_tables(in_size: int, out_size: int):
    """Window start, window length and normalised weights per output index (fp32, like the kernel)."""
    f = np.float32
    scale = f(in_size) / f(out_size)
    support = f(2) * scale if scale >= 1 else f(2)
    invscale = f(1) / scale if scale >= 1 else f(1)
    taps = int(math.ceil(float(support))) * 2 + 1
    xmin = np.zeros(out_size, np.int64)
    xsize = np.zeros(out_size, np.int64)
    w = np.zeros((out_size, taps), np.float32)
    for i in range(out_size):
        center = scale * (f(i) + f(0.5))
        lo = max(int(center - support + f(0.5)), 0)
        n = min(int(center + support + f(0.5)), in_size) - lo
        n = min(max(n, 0), taps)
        j = np.arange(n, dtype=np.float32)
        v = _cubic_aa((j + f(lo) - center + f(0.5)) * invscale)
        total = f(0)
        for t in v:                       # sequential fp32 sum, like the kernel
            total = f(total + t)
        if total != 0:
            v = (v / total).astype(np.float32)
        xmin[i], xsize[i] = lo, n
        w[i, :n] = v
    return xmin, xsize, w


def resize_explicit(x: np.ndarray, size) -> np.ndarray:
    """x [b, c, h, w] fp32 -> [b, c, size[0], size[1]]: width pass, then height pass, fp32 accumulation."""
    b, c, h, w = x.shape
    ho, wo = size
    xmin, xsize, xw = resize_tables(w, wo)
    tmp = np.zeros((b, c, h, wo), np.float32)
    for i in range(wo):
        acc = np.zeros((b, c, h), np.float32)
        for j in range(xsize[i]):
            acc = (acc + x[..., xmin[i] + j] * xw[i, j]).astype(np.float32)
        tmp[..., i] = acc
    ymin, ysize, yw = resize_tables(h, ho)
    out = np.zeros((b, c, ho, wo), np.float32)
    for i in range(ho):
        acc = np.zeros((b, c, wo), np.float32)
        for j in range(ysize[i]):
            acc = (acc + tmp[:, :, ymin[i] + j, :] * yw[i, j]).astype(np.float32)
        out[:, :, i, :] = acc
    return out


# ----------------------------------------------------------------------------------------------- HLS (unpinned)
def rgb_to_hls(image: Tensor) -> Tensor:
    """kornia.color.hls.rgb_to_hls (as called at optex.py:126-127): [b,3,h,w] in [0,1] -> H (radians), L, S."""
    r, g, b = image[:, 0], image[:, 1], image[:, 2]
    maxc, imax = image.max(1)
    minc = image.min(1).values
    l = (maxc + minc) / 2
    d = maxc - minc
    s = torch.where(l < 0.5, d / (maxc + minc), d / (2.0 - (maxc + minc)))
    hi = torch.zeros_like(d)
    hi = torch.where(imax == 0, torch.remainder((g - b) / d, 6), hi)
    hi = torch.where(imax == 1, (b - r) / d + 2, hi)
    hi = torch.where(imax == 2, (r - g) / d + 4, hi)
    h = 2.0 * math.pi * (60.0 * hi) / 360.0
    out = torch.stack([h, l, s], 1)
    return torch.where(torch.isnan(out), torch.zeros_like(out), out)


def hls_to_rgb(image: Tensor) -> Tensor:
    """kornia.color.hls.hls_to_rgb (optex.py:128)."""
    h, l, s = image[:, 0:1], image[:, 1:2], image[:, 2:3]
    off = torch.tensor([0.0, 8.0, 4.0], dtype=image.dtype).view(1, 3, 1, 1)
    h12 = h * (6 / math.pi)
    a = s * torch.min(l, 1.0 - l)
    k = torch.remainder(h12 + off, 12)
    t = torch.min(torch.min(k - 3.0, 9.0 - k), torch.ones_like(k))
    return l - a * torch.max(t, -torch.ones_like(t))


def lightness_transfer(content: Tensor, pastiche: Tensor) -> Tensor:
    """optex.py:126-128: content's hue and saturation, the pastiche's lightness."""
    target = rgb_to_hls(content)
    target[:, 1] = rgb_to_hls(pastiche)[:, 1]
    return hls_to_rgb(target)


# ----------------------------------------------------------------------------------------------- mixing / centring
def mix_layer(A: Tensor, B: Tensor, AtoB: Tensor, BtoA: Tensor, mask_hw: Tensor, alpha: float) -> Tensor:
    """optex.py:197-204 for one layer; mask_hw [mh, mw]; A .. [1, h, w, c]."""
    mix = F.interpolate(mask_hw[None, None], size=A.shape[1:3], mode="nearest").permute(0, 2, 3, 1)
    i = alpha
    return (A * (1 - i) + AtoB * i) * mix + (BtoA * (1 - i) + B * i) * (1 - mix)


def recentre(content_feature: Tensor, style_feature: Tensor) -> Tensor:
    """optex.py:76."""
    return content_feature - content_feature.mean() + torch.mean(style_feature)
