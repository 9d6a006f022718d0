"""Drop-in for the reference's ``util`` module (/root/reference/util.py).

Host arithmetic (schedule, sizes, names) is reproduced bit for bit; the one compute function, `resize`
(util.py:105-106), runs on the B200 (`optex_resize_bicubic_aa`, csrc/image.cu).  Image files are read / written
with PIL on the host like the reference does (util.py:27-30, :45-65) - I/O, not compute.
"""
from __future__ import annotations

from argparse import Namespace
from typing import List, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._runtime import call, f32c, ptr, require_cuda, stream_ptr, workspace


def round32(integer: int) -> int:
    """reference: util.py:93-94."""
    return int(integer + 32 - 1) & -32


def get_size(size: int, scale: float, h: int, w: int, oversize: bool = False) -> Tuple[int, int]:
    """reference: util.py:33-42 (argument names as the reference has them)."""
    ssize = size * scale
    wpercent = ssize / float(h)
    hsize = int((float(w) * float(wpercent)))
    if oversize:
        size = min(int(ssize), h)
        hsize = min(hsize, w)
    return round32(size), round32(hsize)


def get_iters_and_sizes(size: int, iters: int, passes: int, use_multires: bool):
    """reference: util.py:68-86.  Same numpy expressions, so the int32 truncation of the per-layer shares and
    the rounding of the sizes to multiples of 32 come out identically."""
    if use_multires:
        iters_per_pass = np.arange(2 * passes, passes, -1)
        iters_per_pass = iters_per_pass / np.sum(iters_per_pass) * iters
        sizes = np.linspace(256, size, passes)
        sizes = (32 * np.round(sizes / 32)).astype(np.int32).tolist()
    else:
        iters_per_pass = np.ones(passes) * int(iters / passes)
        sizes = [size] * passes
    proportion_per_layer = np.array([64, 128, 256, 512, 512]) + 64
    proportion_per_layer = proportion_per_layer / np.sum(proportion_per_layer)
    its = (iters_per_pass[:, None] * proportion_per_layer[None, :]).astype(np.int32)
    return its.tolist(), sizes


def name(filepath: str) -> str:
    """reference: util.py:89-90."""
    return filepath.split("/")[-1].split(".")[0]


def to_nchw(x: Tensor) -> Tensor:
    """reference: util.py:97-98."""
    return x.permute(0, 3, 1, 2)


def to_nhwc(x: Tensor) -> Tensor:
    """reference: util.py:101-102."""
    return x.permute(0, 2, 3, 1)


def resize(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """reference: util.py:105-106 - antialiased bicubic, align_corners=False.  x [b, c, h, w] on the GPU."""
    dev = require_cuda(x)
    if x.dim() != 4:
        raise ValueError(f"resize expects NCHW [b,c,h,w], got {tuple(x.shape)}")
    b, c, h, w = x.shape
    ho, wo = int(size[0]), int(size[1])
    if min(b, c, h, w, ho, wo) < 1:
        raise ValueError(f"resize: empty input {tuple(x.shape)} or output size {(ho, wo)}")
    xin = f32c(x)
    out = torch.empty(b, c, ho, wo, dtype=torch.float32, device=dev)
    lib = _lib.lib()
    wsb = workspace(dev, lib.optex_resize_workspace_bytes(b * c, h, w, ho, wo))
    with torch.cuda.device(dev):
        call("optex_resize_bicubic_aa", ptr(xin), ptr(out), b * c, h, w, ho, wo, ptr(wsb), wsb.numel(),
             stream_ptr(dev))
    return out.to(x.dtype)


# ----------------------------------------------------------------------------------------------- image files
def load_image(path, size, scale=1, oversize=True, device="cuda", memory_format=torch.contiguous_format) -> Tensor:
    """reference: util.py:27-30 (PIL decode + Lanczos resize on the host, then one upload)."""
    from PIL import Image

    img = Image.open(path).convert(mode="RGB")
    lanczos = getattr(Image, "ANTIALIAS", None) or Image.LANCZOS      # ANTIALIAS was removed in Pillow 10
    img = img.resize(get_size(size, scale, img.size[0], img.size[1], oversize), lanczos)
    arr = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div(255)
    return arr.unsqueeze(0).to(device, memory_format=memory_format)


def load_styles(style_files, size, scale, oversize=False, device="cuda",
                memory_format=torch.contiguous_format) -> List[Tensor]:
    """reference: util.py:13-17."""
    return [load_image(f, size, scale, not oversize, device=device, memory_format=memory_format)
            for f in style_files]


def maybe_load_content(content_file, size, device="cuda", memory_format=torch.contiguous_format):
    """reference: util.py:20-24."""
    if content_file is None:
        return None
    return load_image(content_file, size, oversize=False, device=device, memory_format=memory_format)


def output_name(args: Namespace) -> str:
    """The file stem util.save_image builds (util.py:46-61)."""
    outs = [name(style) for style in args.style]
    if len(args.style) > 1:
        outs += ["blend" + str(args.mixing_alpha)]
    if args.content is not None:
        outs += [name(args.content), "strength" + str(args.content_strength)]
    outs += [args.hist_mode + "hist"]
    if args.no_pca:
        outs += ["no_pca"]
    if args.no_multires:
        outs += ["no_multires"]
    if args.style_scale != 1:
        outs += ["scale" + str(args.style_scale)]
    if args.color_transfer is not None:
        outs += [args.color_transfer]
    outs += [str(args.size)]
    return "_".join(outs)


def save_image(output: Tensor, args: Namespace) -> List[str]:
    """reference: util.py:45-65 (PNG per batch element; torchvision's save_image = clamp, x255 + 0.5, to uint8)."""
    from PIL import Image

    outname = output_name(args)
    paths = []
    for o, out in enumerate(output):
        arr = out.detach().float().cpu().mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
        path = f"{args.output_dir}/{outname}" + (f"_{o + 1}" if len(output) > 1 else "") + ".png"
        Image.fromarray(arr).save(path)
        paths.append(path)
    return paths
