"""`python -m optimaltextures_b200` - the reference's command line (optex.py:209-291) on the B200 path.

Same flags, same defaults, same output file names (util.py:45-65).  Differences, all forced by the hardware target:
  * runs on a B200 only (`--device` is parsed and, like in the reference (optex.py:241 vs :251), not used to pick a CPU);
  * `--no_tf32` selects fp32 FFMA GEMMs, the default is 3xTF32 on the tensor cores (fp32-grade; the reference's default
    is single-pass TF32, available here as OPTEX gemm mode "tf32" through `--gemm tf32`);
  * `--seed` also seeds the rotation stream (the reference's does not: numpy's global RNG, optex.py:149 vs :253);
  * `--script`, `--compile`, `--cudnn_benchmark`, `--memory_format` are accepted and ignored (torch compiler / cuDNN
    front-ends of the reference's library path - there is no such layer here);
  * `--models_dir` (additive): where vgg_normalised_conv{d}_1.pth / feature_invertor_conv{d}_1.pth live (default
    ./models like the reference, or $OPTEX_MODELS_DIR).
"""
from __future__ import annotations

import argparse
import os
from time import time

import torch

from . import optex as _optex
from . import util as _util
from .texture import OptimalTexture


def build_parser() -> argparse.ArgumentParser:
    def required_length(nmin, nmax):
        class RequiredLength(argparse.Action):
            def __call__(self, parser, args, values, option_string=None):
                if not nmin <= len(values) <= nmax:
                    raise argparse.ArgumentTypeError(
                        f'argument "{self.dest}" requires between {nmin} and {nmax} arguments')
                setattr(args, self.dest, values)

        return RequiredLength

    p = argparse.ArgumentParser(prog="python -m optimaltextures_b200")
    p.add_argument("-s", "--style", type=str, nargs="+", action=required_length(1, 2), default=["style/graffiti.jpg"])
    p.add_argument("-c", "--content", type=str, default=None)
    p.add_argument("--batch", type=int, default=1)
    p.add_argument("--size", type=int, default=512)
    p.add_argument("--passes", type=int, default=5)
    p.add_argument("--iters", type=int, default=500)
    p.add_argument("--hist_mode", type=str, choices=["sym", "pca", "chol", "cdf"], default="chol")
    p.add_argument("--color_transfer", type=str, default=None, choices=["lum", "opt"])
    p.add_argument("--content_strength", type=float, default=0.01)
    p.add_argument("--style_scale", type=float, default=1.0)
    p.add_argument("--mixing_alpha", type=float, default=0.5)
    p.add_argument("--no_pca", action="store_true")
    p.add_argument("--no_multires", action="store_true")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--no_tf32", action="store_true")
    p.add_argument("--cudnn_benchmark", action="store_true")
    p.add_argument("--compile", action="store_true")
    p.add_argument("--script", action="store_true")
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--memory_format", type=str, default="contiguous", choices=["contiguous", "channels_last"])
    p.add_argument("--output_dir", type=str, default="output/")
    p.add_argument("--models_dir", type=str, default=None)
    p.add_argument("--gemm", type=str, default=None, choices=["auto", "fp32", "tf32x3", "tf32"])
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("optimaltextures_b200 needs a B200 (sm_100): there is no CPU path - use the reference's optex.py")
    device = "cuda"
    _optex.set_gemm_mode(args.gemm or ("fp32" if args.no_tf32 else "auto"))
    if args.seed is not None:
        torch.manual_seed(args.seed)
        _optex.manual_seed(args.seed)
    models_dir = args.models_dir or os.environ.get("OPTEX_MODELS_DIR") or "models"
    styles = _util.load_styles(args.style, size=args.size, scale=args.style_scale, device=device)
    if len(styles) > 1:
        assert styles[0].shape == styles[1].shape, "Style images must have the same shape"
    content = _util.maybe_load_content(args.content, size=args.size, device=device)
    pastiche = torch.rand(content.shape if content is not None else (args.batch, 3, args.size, args.size)).to(device)
    texturizer = OptimalTexture(size=args.size, iters=args.iters, passes=args.passes, hist_mode=args.hist_mode,
                                color_transfer=args.color_transfer, content_strength=args.content_strength,
                                style_scale=args.style_scale, mixing_alpha=args.mixing_alpha, no_pca=args.no_pca,
                                no_multires=args.no_multires, models_dir=models_dir, device=device)
    torch.cuda.synchronize()
    t = time()
    pastiche = texturizer.forward(pastiche, styles, content, verbose=True)
    torch.cuda.synchronize()
    print("Took:", time() - t)
    os.makedirs(args.output_dir, exist_ok=True)
    _util.save_image(pastiche, args)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
