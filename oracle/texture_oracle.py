"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py) of the synthesis loop around the OT step.

Restates /root/reference/optex.py:15-139 (`OptimalTexture`), :193-206 (`mix_style_features`) and
/root/reference/util.py:33-42,68-106 (`get_size`, `get_iters_and_sizes`, `round32`, `resize`) on top of the other
oracle modules, fp32 on the CPU, calling the same torch ops in the same order as the reference.

What is injected instead of drawn (the reference draws from global RNG state, which no other implementation can
reproduce): `rotation_fn(c, index)` supplies the rotation of the index-th `optimal_transport` call (optex.py:168),
`mask_fn(shape)` the uniform noise of the mixing mask (optex.py:98), `state_dicts` the network weights
(vgg.py:144,162 load them from ./models/*.pth) and `fit_pca_fn` the PCA (default: the reference's, optex.py:180-190;
GPU tests pass the device PCA so that both sides work in the same basis - an SVD basis is only defined up to sign
and to rotations inside near-degenerate subspaces, see tests/test_gpu_texture.py).

Pinned against the real reference by tests/test_oracle_golden.py::test_texture_* (fixture tests/golden/texture.npz,
written by oracle/make_golden.py running optex.OptimalTexture.forward with the same injections).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import image_oracle, ot_oracle, vgg_oracle


# ----------------------------------------------------------------------------------------------- util.py
def round32(integer: int) -> int:
    """util.py:93-94."""
    return int(integer + 32 - 1) & -32


def get_size(size: int, scale: float, h: int, w: int, oversize: bool = False) -> Tuple[int, int]:
    """util.py:33-42 (its `h`, `w` arguments are used as the reference uses them, names included)."""
    ssize = size * scale
    wpercent = ssize / float(h)
    hsize = int((float(w) * float(wpercent)))
    if oversize:
        size = min(int(ssize), h)
        hsize = min(hsize, w)
    return round32(size), round32(hsize)


def get_iters_and_sizes(size: int, iters: int, passes: int, use_multires: bool):
    """util.py:68-86."""
    if use_multires:
        iters_per_pass = np.arange(2 * passes, passes, -1)
        iters_per_pass = iters_per_pass / np.sum(iters_per_pass) * iters
        sizes = np.linspace(256, size, passes)
        sizes = (32 * np.round(sizes / 32)).astype(np.int32)
    else:
        iters_per_pass = np.ones(passes) * int(iters / passes)
        sizes = [size] * passes
    proportion_per_layer = np.array([64, 128, 256, 512, 512]) + 64
    proportion_per_layer = proportion_per_layer / np.sum(proportion_per_layer)
    its = (iters_per_pass[:, None] * proportion_per_layer[None, :]).astype(np.int32)
    return its.tolist(), (sizes.tolist() if hasattr(sizes, "tolist") else list(sizes))


# ----------------------------------------------------------------------------------------------- optex.py
def mix_style_features(style_features: List[Tensor], mixing_mask: Tensor, mixing_alpha: float, hist_mode: str):
    """optex.py:193-206; mixing_mask [1, 1, mh, mw]."""
    out = []
    for sf in style_features:
        A, B = sf[[0]], sf[[1]]
        AtoB = ot_oracle.hist_match_nhwc(A, B, mode=hist_mode)
        BtoA = ot_oracle.hist_match_nhwc(B, A, mode=hist_mode)
        out.append(image_oracle.mix_layer(A, B, AtoB, BtoA, mixing_mask[0, 0], mixing_alpha))
    return out


class OptimalTexture:
    """optex.py:15-139."""

    def __init__(self, state_dicts: Dict[Tuple[str, int], dict], size: int = 512, iters: int = 500, passes: int = 5,
                 hist_mode: str = "chol", color_transfer: Optional[str] = None, content_strength: float = 0.1,
                 style_scale: float = 1, mixing_alpha: float = 0.5, no_pca: bool = False, no_multires: bool = False,
                 rotation_fn: Optional[Callable[[int, int], Tensor]] = None,
                 mask_fn: Optional[Callable[[Tuple[int, int]], Tensor]] = None,
                 fit_pca_fn: Optional[Callable[[Tensor], Tuple[Tensor, Tensor]]] = None):
        self.hist_mode = hist_mode
        self.color_transfer = color_transfer
        self.content_strength = content_strength
        self.style_scale = style_scale
        self.mixing_alpha = mixing_alpha
        self.use_pca = not no_pca
        self.passes = passes
        self.iters_per_pass_and_layer, self.sizes = get_iters_and_sizes(size, iters, passes, not no_multires)
        self.depths = list(range(5, 0, -1))                                            # optex.py:42-43
        self.sd = state_dicts
        self.rotation_fn = rotation_fn
        self.mask_fn = mask_fn or (lambda shape: torch.rand(shape))
        self.fit_pca_fn = fit_pca_fn or ot_oracle.fit_pca
        self.ot_calls = 0
        # set to a list to record every stage as (name, inputs..., output): tests replay the stages one by one on
        # the B200 with the oracle's inputs ("teacher forcing") - the loop as a whole amplifies a 1e-6 input
        # perturbation to 1e-3 .. 1e-2 at the output (random-weight networks), so only per-stage parity is well posed
        self.trace = None

    def _rec(self, *entry):
        if self.trace is not None:
            self.trace.append(tuple(e.clone() if isinstance(e, Tensor) else e for e in entry))
        return entry[-1]

    def encode(self, depth: int, x: Tensor) -> Tensor:
        return self._rec("encode", depth, x, vgg_oracle.encoder_forward(x, self.sd[("encoder", depth)], depth))

    def decode(self, depth: int, f: Tensor) -> Tensor:
        return self._rec("decode", depth, f, vgg_oracle.decoder_forward(f, self.sd[("decoder", depth)], depth))

    def resize(self, x: Tensor, size) -> Tensor:
        return self._rec("resize", x, tuple(size), image_oracle.resize(x, size))

    def optimal_transport(self, pastiche_feature: Tensor, style_feature: Tensor, hist_mode: str) -> Tensor:
        rot = self.rotation_fn(pastiche_feature.shape[-1], self.ot_calls)
        self.ot_calls += 1
        out = ot_oracle.ot_step(pastiche_feature, style_feature, rot, hist_mode)
        return self._rec("ot_step", pastiche_feature, style_feature, rot, hist_mode, out)

    def encode_inputs(self, pastiche: Tensor, styles: List[Tensor], content: Optional[Tensor], size: int):
        """optex.py:45-79."""
        if pastiche.shape[-2] != size and pastiche.shape[-1] != size:
            style_tens = [self.resize(s, get_size(size, self.style_scale, s.shape[2], s.shape[3])) for s in styles]
            if content is not None:
                cont_size = get_size(size, 1.0, content.shape[2], content.shape[3], oversize=True)
                cont_tens = self.resize(content, cont_size)
            else:
                cont_size = (size, size)
                cont_tens = None
            pastiche = self.resize(pastiche, cont_size)
        else:
            style_tens, cont_tens = styles, content
        style_features, style_eigvs, content_features = [], [], []
        for l, depth in enumerate(self.depths):
            style_features.append(torch.cat([self.encode(depth, s) for s in style_tens]))
            eigvecs = None
            if self.use_pca:
                raw = style_features[l]
                style_features[l], eigvecs = self.fit_pca_fn(raw)
                self._rec("fit_pca", raw, eigvecs, style_features[l])
                style_eigvs.append(eigvecs)
            if cont_tens is not None:
                cf = self.encode(depth, cont_tens)
                if self.use_pca:
                    cf = self._rec("project", cf, eigvecs, False, cf @ eigvecs)
                content_features.append(self._rec("recentre", cf, style_features[l],
                                                  image_oracle.recentre(cf, style_features[l])))
        return pastiche, style_features, style_eigvs, content_features

    def forward(self, pastiche: Tensor, styles: List[Tensor], content: Optional[Tensor] = None) -> Tensor:
        """optex.py:81-139."""
        for p in range(self.passes):
            pastiche, style_features, style_eigvs, content_features = self.encode_inputs(
                pastiche, styles, content, self.sizes[p])
            if len(styles) > 1:
                mask = torch.ceil(self.mask_fn(tuple(style_features[1].shape[1:3])) - self.mixing_alpha)[None, None]
                mixed = mix_style_features(style_features, mask, self.mixing_alpha, self.hist_mode)
                for sf, mx in zip(style_features, mixed):
                    self._rec("mix", sf, mask, self.mixing_alpha, self.hist_mode, mx)
                style_features = mixed
            for l, depth in enumerate(self.depths):
                f = self.encode(depth, pastiche)
                if self.use_pca:
                    f = self._rec("project", f, style_eigvs[l], False, f @ style_eigvs[l])
                f_in, first = f, self.ot_calls
                blend = len(content_features) > 0 and l <= 2
                strength = self.content_strength / 2 ** (4 - l)
                for _ in range(self.iters_per_pass_and_layer[p][l - 1]):     # the reference's [l - 1], optex.py:112
                    f = self.optimal_transport(f, style_features[l], self.hist_mode)
                    if blend:
                        f = f + strength * (content_features[l] - f)
                if self.ot_calls > first:
                    self._rec("ot_loop", f_in, style_features[l], first, self.ot_calls - first, self.hist_mode,
                              content_features[l] if blend else None, strength if blend else 0.0, f)
                if self.use_pca:
                    f = self._rec("project", f, style_eigvs[l], True, f @ style_eigvs[l].T)
                pastiche = self.decode(depth, f)
        if self.color_transfer is not None:
            assert content is not None, "Color transfer requires content image"
            target = self._rec("lightness", content, pastiche, image_oracle.lightness_transfer(content, pastiche))
            if self.color_transfer == "opt":
                pastiche, target = pastiche.permute(0, 2, 3, 1), target.permute(0, 2, 3, 1)
                for _ in range(3):
                    pastiche = self.optimal_transport(pastiche, target, "cdf")
                pastiche = pastiche.permute(0, 3, 1, 2)
            elif self.color_transfer == "lum":
                pastiche = target
        return pastiche
