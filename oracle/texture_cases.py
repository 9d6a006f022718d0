"""Seeded cases of the synthesis-loop parity tests (TEST INFRASTRUCTURE): shared by oracle/make_golden.py (which
runs the real reference on them) and tests/ (which run the oracle and the B200 path on them)."""
import torch

from . import vgg_oracle


TEXTURE_CASES = {
    # name: OptimalTexture kwargs, style shapes, content shape (or None), pastiche shape.  Multires only: the
    # reference's no_multires path raises (util.py:86 calls .tolist() on a list).  Pass 0 always runs at 256^2.
    "synth_pca": (dict(size=64, iters=24, passes=2, hist_mode="pca"), [(1, 3, 80, 64)], None, (1, 3, 64, 64)),
    "mix_content_chol_opt": (dict(size=96, iters=20, passes=2, hist_mode="chol", color_transfer="opt",
                                  content_strength=0.2, mixing_alpha=0.4),
                             [(1, 3, 64, 64), (1, 3, 64, 64)], (1, 3, 96, 96), (1, 3, 96, 96)),
    "nopca_cdf_lum": (dict(size=32, iters=12, passes=2, hist_mode="cdf", no_pca=True, color_transfer="lum",
                           content_strength=0.1), [(1, 3, 40, 48)], (1, 3, 32, 32), (1, 3, 32, 32)),
}


def texture_inputs(name):
    """Seeded inputs of a TEXTURE_CASES entry (shared with tests/)."""
    import zlib

    kwargs, style_shapes, content_shape, pastiche_shape = TEXTURE_CASES[name]
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    styles = [torch.rand(*s, generator=g) for s in style_shapes]
    content = torch.rand(*content_shape, generator=g) if content_shape else None
    pastiche = torch.rand(*pastiche_shape, generator=g)
    return kwargs, styles, content, pastiche


def texture_rotation(c, index):
    """The injected rotation stream: scipy's construction (optex.py:149) from an explicit per-call seed."""
    from scipy.stats import special_ortho_group

    if c == 1:
        return torch.ones(1, 1, dtype=torch.float64)
    return torch.tensor(special_ortho_group.rvs(c, random_state=5000 + index))


def state_dicts(seed: int = 0):
    """The seeded random weights the golden run used in place of ./models/*.pth."""
    sd = {}
    for d in range(1, 6):
        sd[("encoder", d)] = vgg_oracle.random_state_dict("encoder", d, seed)
        sd[("decoder", d)] = vgg_oracle.random_state_dict("decoder", d, seed)
    return sd
