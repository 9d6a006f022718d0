"""Drop-in for the reference's ``vgg`` module (/root/reference/vgg.py) on the B200 kernels.

    Encoder(depth)(image NCHW) -> NHWC features      vgg.py:138-153   (called at optex.py:62-63, 107)
    Decoder(depth)(features NHWC) -> NCHW image      vgg.py:156-171   (called at optex.py:122)

Each 3x3 conv layer is one `optex_conv3x3` call (csrc/vgg.cu): gather with the reflection padding / ceil-mode
max-pool / nearest up-sampling folded into its indices, then the tcgen05 GEMM with bias + ReLU in the epilogue,
NHWC in and out.  The encoder's leading 1x1 colour conv (vgg.py:16) is folded into conv1_1's weights at load time
(a pointwise map commutes with the reflection padding).

Weights: the reference's state_dicts (`vgg_normalised_conv{d}_1.pth`, `feature_invertor_conv{d}_1.pth`), passed as
`state_dict=` or found in `models_dir=` / $OPTEX_MODELS_DIR.  They are not part of this repository.

Additive: `Encoder.forward_all(x)` returns the features of conv1_1 .. conv{depth}_1 from ONE pass - the reference's
five encoder files are prefix-identical, so `Encoder(5).forward_all` replaces five encodes of the
same image (optex.py:62-63 runs all five on every style image).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from torch import Tensor

from . import _lib
from ._runtime import call, f32c, ptr, require_cuda, stream_ptr, workspace

NONE, POOL, UP = 0, 1, 2

# (pre-op, c_in, c_out, relu) of every 3x3 conv, in execution order (vgg.py:14-75)
_ENCODER = [
    (NONE, 3, 64, True),
    (NONE, 64, 64, True), (POOL, 64, 128, True),
    (NONE, 128, 128, True), (POOL, 128, 256, True),
    (NONE, 256, 256, True), (NONE, 256, 256, True), (NONE, 256, 256, True), (POOL, 256, 512, True),
    (NONE, 512, 512, True), (NONE, 512, 512, True), (NONE, 512, 512, True), (POOL, 512, 512, True),
]
_ENCODER_END = {1: 1, 2: 3, 3: 5, 4: 9, 5: 13}
# decoder blocks, deepest first (vgg.py:78-136); Decoder(d) runs the last d
_DECODER = [
    [(NONE, 512, 512, True), (UP, 512, 512, True), (NONE, 512, 512, True), (NONE, 512, 512, True)],
    [(NONE, 512, 256, True), (UP, 256, 256, True), (NONE, 256, 256, True), (NONE, 256, 256, True)],
    [(NONE, 256, 128, True), (UP, 128, 128, True)],
    [(NONE, 128, 64, True), (UP, 64, 64, True)],
    [(NONE, 64, 3, False)],
]


def _load_state_dict(state_dict, models_dir, fname):
    if state_dict is not None:
        return state_dict
    models_dir = models_dir or os.environ.get("OPTEX_MODELS_DIR")
    if not models_dir:
        raise RuntimeError(f"no weights: pass state_dict= or models_dir= (or set OPTEX_MODELS_DIR) containing {fname}")
    return torch.load(os.path.join(models_dir, fname), map_location="cpu")


def random_state_dicts(seed: int = 0):
    """Random-init weights of the reference's architecture, keyed ("encoder"|"decoder", depth) in its state_dict
    order (vgg.py:14-136): what bench.py and the CLI's --random_weights use where ./models/*.pth are absent.  The
    encoder keeps the 1x1 colour conv's scale (x255, BGR means) so the activations have the reference's range."""
    out = {}
    for depth in range(1, 6):
        for kind in ("encoder", "decoder"):
            g = torch.Generator().manual_seed(seed)
            sd, idx = {}, 0
            if kind == "encoder":
                specs = _ENCODER[:_ENCODER_END[depth]]
                sd["0.weight"] = torch.eye(3).reshape(3, 3, 1, 1) * 255.0 + torch.randn(3, 3, 1, 1, generator=g)
                sd["0.bias"] = torch.tensor([-103.9, -116.8, -123.7])
                idx = 1
            else:
                specs = [spec for block in _DECODER[5 - depth:] for spec in block]
            for _, c_in, c_out, _ in specs:
                scale = (2.0 / (9 * c_in)) ** 0.5 * (0.02 if (kind == "encoder" and c_in == 3) else 1.0)
                sd[f"{idx}.weight"] = torch.randn(c_out, c_in, 3, 3, generator=g) * scale
                sd[f"{idx}.bias"] = torch.randn(c_out, generator=g) * 0.05
                idx += 1
            out[(kind, depth)] = sd
    return out


def _pairs(state_dict):
    vals = list(state_dict.values())
    if len(vals) % 2:
        raise ValueError("state_dict must alternate weight, bias")
    return [(vals[i].detach().double().cpu(), vals[i + 1].detach().double().cpu()) for i in range(0, len(vals), 2)]


def _pack(w: Tensor, b: Tensor, c_out_pad: int, device):
    """torch [co, ci, 3, 3] -> the GEMM operand [co_pad, kp] of optex_conv3x3 (see include/optex_b200.h)."""
    co, ci = w.shape[:2]
    kp = _lib.lib().optex_conv3x3_packed_k(ci)
    m = torch.zeros(c_out_pad, kp, dtype=torch.float64)
    m[:co, :9 * ci] = w.permute(0, 2, 3, 1).reshape(co, 9 * ci)
    bb = torch.zeros(c_out_pad, dtype=torch.float64)
    bb[:co] = b
    return m.float().contiguous().to(device), bb.float().contiguous().to(device)


class _Layer:
    __slots__ = ("pre", "cin", "cout", "cout_pad", "relu", "w", "b")

    def __init__(self, spec, w, b, device):
        self.pre, self.cin, self.cout, self.relu = spec
        if tuple(w.shape) != (self.cout, self.cin, 3, 3):
            raise ValueError(f"weight {tuple(w.shape)} does not fit a {self.cin}->{self.cout} 3x3 conv")
        # narrow outputs (the decoder's 64 -> 3) are padded to 32 columns so the layer stays on the tensor cores
        self.cout_pad = self.cout if self.cout % 8 == 0 else (self.cout + 31) // 32 * 32
        self.w, self.b = _pack(w, b, self.cout_pad, device)

    def run(self, x: Tensor, src_nchw: bool = False) -> Tensor:
        dev = x.device
        if src_nchw:
            b, c, hs, ws = x.shape
        else:
            b, hs, ws, c = x.shape
        if c != self.cin:
            raise ValueError(f"layer expects {self.cin} channels, got {c}")
        h, w = ((hs + 1) // 2, (ws + 1) // 2) if self.pre == POOL else ((2 * hs, 2 * ws) if self.pre == UP else (hs, ws))
        out = torch.empty(b, h, w, self.cout_pad, dtype=torch.float32, device=dev)
        lib = _lib.lib()
        wsb = workspace(dev, lib.optex_conv3x3_workspace_bytes(b, hs, ws, self.cin, self.cout_pad, self.pre))
        with torch.cuda.device(dev):
            call("optex_conv3x3", ptr(x), 1 if src_nchw else 0, b, hs, ws, self.cin, ptr(self.w), ptr(self.b),
                 self.cout_pad, self.pre, 1 if self.relu else 0, ptr(out), self.cout_pad, ptr(wsb), wsb.numel(),
                 stream_ptr(dev))
        return out


class Encoder:
    """reference: vgg.py:138-153 (`Encoder(depth)`; weights vgg_normalised_conv{depth}_1.pth)."""

    def __init__(self, depth: int, state_dict=None, models_dir: Optional[str] = None, device="cuda"):
        assert isinstance(depth, int) and 1 <= depth <= 5
        self.depth = depth
        self.device = torch.device(device)
        wb = _pairs(_load_state_dict(state_dict, models_dir, f"vgg_normalised_conv{depth}_1.pth"))
        specs = _ENCODER[:_ENCODER_END[depth]]
        if len(wb) != len(specs) + 1:
            raise ValueError(f"Encoder({depth}) needs {len(specs) + 1} weight/bias pairs, got {len(wb)}")
        (w0, b0), (w1, b1) = wb[0], wb[1]                      # fold the 1x1 colour conv (vgg.py:16) into conv1_1
        w0 = w0.reshape(3, 3)
        w1f = torch.einsum("ojyx,ji->oiyx", w1, w0)
        b1f = b1 + torch.einsum("ojyx,j->o", w1, b0)
        wb = [(w1f, b1f)] + wb[2:]
        self.layers = [_Layer(s, w, b, self.device) for s, (w, b) in zip(specs, wb)]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def forward_all(self, x: Tensor) -> List[Tensor]:
        """NHWC features of conv1_1 .. conv{depth}_1 from one pass."""
        require_cuda(x)
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected an NCHW image [b,3,H,W], got {tuple(x.shape)}")
        cur = f32c(x)
        ends = set(_ENCODER_END[d] for d in range(1, self.depth + 1))
        outs = []
        for i, layer in enumerate(self.layers, start=1):
            cur = layer.run(cur, src_nchw=(i == 1))
            if i in ends:
                outs.append(cur)
        return outs

    def forward(self, x: Tensor) -> Tensor:
        return self.forward_all(x)[-1].to(x.dtype)

    __call__ = forward


class Decoder:
    """reference: vgg.py:156-171 (`Decoder(depth)`; weights feature_invertor_conv{depth}_1.pth)."""

    def __init__(self, depth: int, state_dict=None, models_dir: Optional[str] = None, device="cuda"):
        assert isinstance(depth, int) and 1 <= depth <= 5
        self.depth = depth
        self.device = torch.device(device)
        wb = _pairs(_load_state_dict(state_dict, models_dir, f"feature_invertor_conv{depth}_1.pth"))
        specs = [c for block in _DECODER[-depth:] for c in block]
        if len(wb) != len(specs):
            raise ValueError(f"Decoder({depth}) needs {len(specs)} weight/bias pairs, got {len(wb)}")
        self.layers = [_Layer(s, w, b, self.device) for s, (w, b) in zip(specs, wb)]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def forward(self, x: Tensor) -> Tensor:
        dev = require_cuda(x)
        if x.dim() != 4:
            raise ValueError(f"expected NHWC features [b,h,w,c], got {tuple(x.shape)}")
        cur = f32c(x)
        for layer in self.layers:
            cur = layer.run(cur)
        b, h, w, cp = cur.shape
        out = torch.empty(b, 3, h, w, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("optex_nhwc_to_nchw", ptr(cur), ptr(out), b, h * w, cp, 3, stream_ptr(dev))
        return out.to(x.dtype)

    __call__ = forward


def conv3x3(x: Tensor, weight: Tensor, bias: Optional[Tensor], pre_op: int = NONE, relu: bool = True,
            src_nchw: bool = False) -> Tensor:
    """One layer with torch-layout weights [co, ci, 3, 3] (packs them on every call: tests and one-off use)."""
    dev = require_cuda(x)
    co, ci = weight.shape[:2]
    b = bias if bias is not None else torch.zeros(co)
    layer = _Layer((pre_op, ci, co, relu), weight.detach().double().cpu(), b.detach().double().cpu(), dev)
    return layer.run(f32c(x), src_nchw=src_nchw)[..., :co]
