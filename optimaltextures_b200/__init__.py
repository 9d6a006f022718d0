"""optimaltextures_b200 - the sliced-OT hot path of JCBrouwer/OptimalTextures on B200 (sm_100a).

Host-side mirror of the reference's call surface over liboptex_b200.so (C-ABI, include/optex_b200.h).
CUDA only: importing works anywhere, every compute call needs a B200 and the built library.
"""
from . import util, vgg  # noqa: F401
from .histmatch import cdf_match, hist_match, interp, sort_match  # noqa: F401
from .optex import (fit_pca, install, manual_seed, pca_project, optimal_transport, optimal_transport_host, ot_loop, prepared_rotation,  # noqa: F401
                    random_rotation, random_rotations, rotate_forward, rotate_inverse, set_gemm_mode,
                    set_host_style, set_rotation_precision)
from .texture import (OptimalTexture, hls_to_rgb, lightness_transfer, mix_style_features, recentre,  # noqa: F401
                      rgb_to_hls)
from .util import get_iters_and_sizes, get_size, resize  # noqa: F401

__all__ = ["hist_match", "cdf_match", "sort_match", "interp", "optimal_transport", "optimal_transport_host", "set_host_style",
           "ot_loop", "random_rotation", "random_rotations", "rotate_forward", "rotate_inverse", "manual_seed", "set_gemm_mode",
           "set_rotation_precision", "prepared_rotation", "install", "fit_pca", "pca_project", "OptimalTexture",
           "mix_style_features", "lightness_transfer", "rgb_to_hls", "hls_to_rgb", "recentre", "resize", "get_size",
           "get_iters_and_sizes", "util", "vgg"]
