"""Import the UNMODIFIED reference (baseline/_ref/, else /root/reference) for the reference arm of bench.py, the
cpu_baseline leg and tests/.  TEST / MEASUREMENT INFRASTRUCTURE - the product never imports this.

Two shims are needed to import it in this image, neither touches its arithmetic on the hot path:
  * `kornia` is not installed; optex.py:5 imports kornia.color.hls, used only by colour transfer (optex.py:126-128).
    A stub module is registered; its two functions come from oracle/image_oracle.py (a restatement of the published
    HLS formulas - parity unpinned, only reached with --color_transfer).
  * Pillow >= 10 dropped Image.ANTIALIAS (util.py:29): aliased to Image.LANCZOS, the same filter.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CANDIDATES = [os.path.join(HERE, "_ref"), "/root/reference"]


def path():
    for p in CANDIDATES:
        if os.path.exists(os.path.join(p, "optex.py")) and os.path.exists(os.path.join(p, "histmatch.py")):
            return p
    return None


def available() -> bool:
    return path() is not None


def has_weights() -> bool:
    p = path()
    return p is not None and os.path.exists(os.path.join(p, "models", "vgg_normalised_conv5_1.pth"))


_loaded = None


def load():
    """-> namespace with .optex .histmatch .util .vgg .path  (the reference's own modules, unmodified)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    p = path()
    if p is None:
        raise RuntimeError("the reference is not staged: run `python baseline/stage_reference.py` in the build "
                           "container (needs /root/reference)")
    if "kornia" not in sys.modules:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        kornia = types.ModuleType("kornia")
        color = types.ModuleType("kornia.color")
        hls = types.ModuleType("kornia.color.hls")

        def _rgb_to_hls(x):
            from oracle import image_oracle
            return image_oracle.rgb_to_hls(x)

        def _hls_to_rgb(x):
            from oracle import image_oracle
            return image_oracle.hls_to_rgb(x)

        hls.rgb_to_hls, hls.hls_to_rgb = _rgb_to_hls, _hls_to_rgb
        kornia.color, color.hls = color, hls
        sys.modules.update({"kornia": kornia, "kornia.color": color, "kornia.color.hls": hls})
    from PIL import Image

    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    # the reference's modules import each other by bare name (optex.py:9-11) and vgg.py opens ./models/... relative to
    # the working directory (vgg.py:144,162): import with its directory first on sys.path, from inside it
    for name in ("optex", "histmatch", "util", "vgg"):
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(p):
            raise RuntimeError(f"a different module named {name!r} is already imported ({m.__file__})")
    sys.path.insert(0, p)
    cwd = os.getcwd()
    os.chdir(p)
    try:
        import histmatch
        import optex
        import util
        import vgg
    finally:
        os.chdir(cwd)
        sys.path.remove(p)
    _loaded = types.SimpleNamespace(optex=optex, histmatch=histmatch, util=util, vgg=vgg, path=p)
    return _loaded


class in_reference_dir:
    """`with in_reference_dir():` - the reference resolves ./models/*.pth and style/ paths against the cwd."""

    def __enter__(self):
        self.cwd = os.getcwd()
        os.chdir(load().path)
        return self

    def __exit__(self, *exc):
        os.chdir(self.cwd)
        return False
