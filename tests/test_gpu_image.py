"""GPU parity: the image-space glue of a synthesis pass through the C-ABI (csrc/image.cu) vs oracle/image_oracle.py.

Stated tolerances (floating point):
  * resize (util.py:105-106): |out - torch CPU interpolate| <= 1e-5 (inputs in [0, 1]; both sides are fp32 sums of
    <= 2 x 35 products, the explicit arithmetic is pinned to torch at 4e-6 in tests/test_texture_host.py);
  * HLS (optex.py:126-128; kornia absent -> parity UNPINNED, oracle restates its formulas): 1e-5 on values in [0, 1]
    away from the hue wrap-around; rgb -> hls -> rgb round trip 1e-5;
  * mix_style_features blend (optex.py:203) and the content re-centring (optex.py:76): 1e-6 * scale (same fp32
    expression; the scalar means are accumulated in FP64 on the device, in fp32 pairwise by torch)."""
import numpy as np
import pytest
import torch

from oracle import image_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import optimaltextures_b200 as ob

    return ob


RESIZES = [(1, 3, 64, 48, 32, 32), (2, 3, 37, 53, 64, 96), (1, 3, 256, 256, 448, 448), (1, 3, 416, 416, 256, 224),
           (1, 1, 100, 80, 33, 95), (1, 3, 32, 32, 32, 64), (1, 3, 5, 7, 1, 1), (1, 3, 1024, 1024, 256, 256),
           (1, 3, 256, 256, 1024, 1024), (1, 3, 64, 64, 64, 64)]


@pytest.mark.parametrize("b,c,h,w,ho,wo", RESIZES)
def test_resize_matches_torch_cpu(ob, b, c, h, w, ho, wo):
    from optimaltextures_b200 import util

    x = torch.rand(b, c, h, w, generator=torch.Generator().manual_seed(h * 1000 + w))
    want = image_oracle.resize(x, (ho, wo))
    got = util.resize(x.cuda(), (ho, wo)).cpu()
    assert got.shape == want.shape
    err = float((got - want).abs().max())
    assert err <= 1e-5, f"max |err| {err:.3e}"


def test_resize_errors(ob):
    from optimaltextures_b200 import util

    with pytest.raises(ValueError):
        util.resize(torch.zeros(3, 8, 8, device="cuda"), (4, 4))
    with pytest.raises(ValueError):
        util.resize(torch.zeros(1, 3, 8, 8, device="cuda"), (0, 4))


def _images(seed, shape=(2, 3, 40, 56)):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(*shape, generator=g)
    x[0, :, 0, :8] = torch.tensor([[1.0, 0, 0, 0.5, 0, 1, 0.2, 0.7], [0, 1.0, 0, 0.5, 0, 1, 0.4, 0.7],
                                   [0, 0, 1.0, 0.5, 0, 1, 0.6, 0.1]])      # primaries, grey, black, white
    return x


def test_rgb_to_hls_and_back(ob):
    from optimaltextures_b200 import texture

    x = _images(1)
    hls = texture.rgb_to_hls(x.cuda()).cpu()
    ref = image_oracle.rgb_to_hls(x)
    assert not torch.isnan(hls).any()
    # hue is an angle: compare on the circle
    dh = torch.remainder(hls[:, 0] - ref[:, 0] + np.pi, 2 * np.pi) - np.pi
    assert float(dh.abs().max()) <= 2e-5
    assert float((hls[:, 1:] - ref[:, 1:]).abs().max()) <= 1e-5
    back = texture.hls_to_rgb(hls.cuda()).cpu()
    assert float((back - x).abs().max()) <= 1e-5
    assert float((texture.hls_to_rgb(ref.cuda()).cpu() - image_oracle.hls_to_rgb(ref)).abs().max()) <= 1e-5


def test_lightness_transfer(ob):
    from optimaltextures_b200 import texture

    content, pastiche = _images(2), _images(3)
    got = texture.lightness_transfer(content.cuda(), pastiche.cuda()).cpu()
    want = image_oracle.lightness_transfer(content, pastiche)
    assert float((got - want).abs().max()) <= 1e-5
    # the pastiche may leave [0, 1] (a decoder output): same formulas, still finite and equal
    wild = pastiche * 1.6 - 0.3
    got = texture.lightness_transfer(content.cuda(), wild.cuda()).cpu()
    want = image_oracle.lightness_transfer(content, wild)
    assert float((got - want).abs().max()) <= 2e-5
    with pytest.raises(ValueError):
        texture.lightness_transfer(content.cuda(), pastiche[:, :, :8].cuda())


@pytest.mark.parametrize("h,w,c,mh,mw", [(16, 16, 8, 16, 16), (32, 24, 5, 8, 6), (7, 9, 3, 16, 16), (64, 64, 64, 32, 32)])
def test_mix_features(ob, h, w, c, mh, mw):
    from optimaltextures_b200 import _lib
    from optimaltextures_b200._runtime import call, ptr, stream_ptr

    g = torch.Generator().manual_seed(h * w + c)
    A, B, AtoB, BtoA = (torch.randn(1, h, w, c, generator=g) for _ in range(4))
    mask = torch.ceil(torch.rand(mh, mw, generator=g) - 0.4)
    want = image_oracle.mix_layer(A, B, AtoB, BtoA, mask, 0.3)
    dev = [t.cuda() for t in (A, B, AtoB, BtoA, mask)]
    out = torch.empty_like(dev[0])
    call("optex_mix_features", *(ptr(t) for t in dev), ptr(out), h, w, c, mh, mw, 0.3, stream_ptr(out.device))
    assert float((out.cpu() - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))


def test_mix_style_features_end_to_end(ob):
    """optex.py:193-206 with the two hist_match calls on the device (mode chol), vs the oracle's."""
    from optimaltextures_b200 import texture
    from oracle import texture_oracle

    g = torch.Generator().manual_seed(5)
    feats = [torch.relu(torch.randn(2, h, h, c, generator=g)) for h, c in ((4, 48), (8, 32), (16, 16))]
    mask = torch.ceil(torch.rand(8, 8, generator=g) - 0.5)[None, None]
    want = texture_oracle.mix_style_features([f.clone() for f in feats], mask, 0.5, "chol")
    got = texture.mix_style_features([f.cuda() for f in feats], mask.cuda(), 0.5, "chol")
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert float((a.cpu() - b).abs().max()) <= 5e-4 * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("n,m", [(1, 1), (1000, 77), (1 << 20, 300000)])
def test_recentre(ob, n, m):
    from optimaltextures_b200 import texture

    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g).reshape(1, 1, n, 1) * 3 + 1
    s = torch.rand(m, generator=g).reshape(1, 1, m, 1) * 5
    want = image_oracle.recentre(x, s)
    got = texture.recentre(x.cuda(), s.cuda()).cpu()
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))
