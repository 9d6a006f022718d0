"""CPU oracle (test infrastructure, never the product path) for the VGG-19 encoder / feature-inverter stacks.

Restates /root/reference/vgg.py in torch functional ops, fp32 on the CPU:
  encoder (vgg.py:14-75):  Conv2d(3,3,1x1) then [ReflectionPad2d(1) -> Conv2d 3x3 -> ReLU] layers with
                           MaxPool2d((2,2),(2,2),ceil_mode=True) in front of conv2_1, conv3_1, conv4_1, conv5_1;
                           Encoder.forward returns NHWC (vgg.py:152-153)
  decoder (vgg.py:78-136): the mirrored stacks with UpsamplingNearest2d(scale_factor=2); the last conv (64 -> 3)
                           has no ReLU; Decoder.forward takes NHWC, returns NCHW (vgg.py:170-171)
Weights come as the reference's state_dict (an ordered mapping whose values alternate weight, bias in layer
order - the key names, Sequential indices, are not interpreted).

Pinned against the real reference in tests/test_oracle_golden.py::test_vgg_* (fixture tests/golden/vgg.npz written by
oracle/make_golden.py from vgg.Encoder(2) / vgg.Decoder(2) with the reference's own .pth weights).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

NONE, POOL, UP = 0, 1, 2

# (pre-op, c_in, c_out, relu) per 3x3 conv; ENCODER_DEPTH_END[d] = number of convs of Encoder(d)
ENCODER_CONVS: List[Tuple[int, int, int, bool]] = [
    (NONE, 3, 64, True),                                                                      # conv1_1
    (NONE, 64, 64, True), (POOL, 64, 128, True),                                              # conv2_1
    (NONE, 128, 128, True), (POOL, 128, 256, True),                                           # conv3_1
    (NONE, 256, 256, True), (NONE, 256, 256, True), (NONE, 256, 256, True), (POOL, 256, 512, True),   # conv4_1
    (NONE, 512, 512, True), (NONE, 512, 512, True), (NONE, 512, 512, True), (POOL, 512, 512, True),   # conv5_1
]
ENCODER_DEPTH_END = {1: 1, 2: 3, 3: 5, 4: 9, 5: 13}

# decoder blocks, deepest first (feature_invertor, vgg.py:78-136); Decoder(d) uses the LAST d blocks
DECODER_BLOCKS: List[List[Tuple[int, int, int, bool]]] = [
    [(NONE, 512, 512, True), (UP, 512, 512, True), (NONE, 512, 512, True), (NONE, 512, 512, True)],   # from conv5_1
    [(NONE, 512, 256, True), (UP, 256, 256, True), (NONE, 256, 256, True), (NONE, 256, 256, True)],   # from conv4_1
    [(NONE, 256, 128, True), (UP, 128, 128, True)],                                                   # from conv3_1
    [(NONE, 128, 64, True), (UP, 64, 64, True)],                                                      # from conv2_1
    [(NONE, 64, 3, False)],                                                                           # from conv1_1
]


def encoder_convs(depth: int):
    return ENCODER_CONVS[:ENCODER_DEPTH_END[depth]]


def decoder_convs(depth: int):
    return [c for block in DECODER_BLOCKS[-depth:] for c in block]


def pairs(state_dict) -> List[Tuple[Tensor, Tensor]]:
    vals = [v for v in state_dict.values()]
    assert len(vals) % 2 == 0
    return [(vals[i].float(), vals[i + 1].float()) for i in range(0, len(vals), 2)]


def _layer(x: Tensor, pre: int, w: Tensor, b: Tensor, relu: bool) -> Tensor:
    if pre == POOL:
        x = F.max_pool2d(x, (2, 2), (2, 2), (0, 0), ceil_mode=True)
    elif pre == UP:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b)
    return torch.relu(x) if relu else x


def encoder_forward(x_nchw: Tensor, state_dict, depth: int, all_depths: bool = False):
    """vgg.py:138-153.  x [b,3,H,W] -> NHWC features of conv{depth}_1 (all_depths: list for 1..depth)."""
    wb = pairs(state_dict)
    convs = encoder_convs(depth)
    assert len(wb) == len(convs) + 1, f"{len(wb)} weight pairs for Encoder({depth})"
    x = F.conv2d(x_nchw.float(), wb[0][0], wb[0][1])          # the 1x1 colour conv, vgg.py:16
    outs = []
    ends = {ENCODER_DEPTH_END[d]: d for d in range(1, depth + 1)}
    for i, ((pre, _, _, relu), (w, b)) in enumerate(zip(convs, wb[1:]), start=1):
        x = _layer(x, pre, w, b, relu)
        if i in ends:
            outs.append(x.permute(0, 2, 3, 1).contiguous())
    return outs if all_depths else outs[-1]


def decoder_forward(x_nhwc: Tensor, state_dict, depth: int) -> Tensor:
    """vgg.py:156-171.  NHWC features of conv{depth}_1 -> image [b,3,H,W]."""
    wb = pairs(state_dict)
    convs = decoder_convs(depth)
    assert len(wb) == len(convs), f"{len(wb)} weight pairs for Decoder({depth})"
    x = x_nhwc.float().permute(0, 3, 1, 2)
    for (pre, _, _, relu), (w, b) in zip(convs, wb):
        x = _layer(x, pre, w, b, relu)
    return x


def random_state_dict(kind: str, depth: int, seed: int = 0):
    """He-initialised weights in the reference's state_dict order (tests on boxes without the .pth files)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    convs = encoder_convs(depth) if kind == "encoder" else decoder_convs(depth)
    idx = 0
    if kind == "encoder":
        sd["0.weight"] = torch.eye(3).reshape(3, 3, 1, 1) * 255.0 + torch.randn(3, 3, 1, 1, generator=g)
        sd["0.bias"] = torch.tensor([-103.9, -116.8, -123.7])
        idx = 1
    for _, cin, cout, _ in convs:
        scale = (2.0 / (9 * cin)) ** 0.5 * (0.02 if (kind == "encoder" and cin == 3) else 1.0)
        sd[f"{idx}.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * scale
        sd[f"{idx}.bias"] = torch.randn(cout, generator=g) * 0.05
        idx += 1
    return sd
