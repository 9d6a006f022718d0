"""GPU tests against the REAL reference staged under baseline/_ref/ (its four .py files, its trained models/*.pth and
its bundled style/ content/ images - baseline/stage_reference.py; git-ignored, travels to the GPU box):

  (i)   Encoder / Decoder of this package with the reference's real weights vs the reference's own torch modules
        (vgg.py:138-171) on CPU, at the bundled image;
  (ii)  the reference's own `OptimalTexture.forward` (optex.py:81-139) as the CALLER after `ob.install(optex)`: its
        loop drives optimal_transport / hist_match / fit_pca / resize of this package on the B200 - compared with
        the same forward run by the untouched reference on the CPU, same pastiche noise;
  (iii) one OT step on real conv4_1 features of the bundled style image, reference function vs B200.

Skipped (not passed) when baseline/_ref is absent.
"""
import os

import numpy as np
import pytest
import torch

from baseline import reference

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (reference.available() and reference.has_weights()),
                                 reason="baseline/_ref not staged (python baseline/stage_reference.py)")]


@pytest.fixture(scope="module")
def ob():
    import optimaltextures_b200 as ob

    return ob


@pytest.fixture(scope="module")
def ref():
    torch.backends.cuda.matmul.allow_tf32 = False      # optex.py:248-249 sets these from --no_tf32; parity wants fp32
    torch.backends.cudnn.allow_tf32 = False
    return reference.load()


def load_style(ref, name, size):
    return ref.util.load_styles([os.path.join(ref.path, "style", name)], size=size, scale=1.0)[0]


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 5])
def test_vgg_real_weights_vs_reference_modules(ob, ref, depth):
    from optimaltextures_b200 import vgg

    img = load_style(ref, "graffiti.jpg", 128)                       # [1, 3, 192, 128]
    with reference.in_reference_dir(), torch.inference_mode():
        enc_ref, dec_ref = ref.vgg.Encoder(depth), ref.vgg.Decoder(depth)
        f_ref = enc_ref(img)
        back_ref = dec_ref(f_ref)
    models = os.path.join(ref.path, "models")
    f = vgg.Encoder(depth, models_dir=models)(img.cuda()).cpu()
    assert f.shape == f_ref.shape
    scale = float(f_ref.abs().max())
    assert float((f - f_ref).abs().max()) <= 1e-3 * scale, f"Encoder({depth}) off by {float((f - f_ref).abs().max())}"
    back = vgg.Decoder(depth, models_dir=models)(f_ref.cuda()).cpu()
    assert back.shape == back_ref.shape
    assert float((back - back_ref).abs().max()) <= 1e-3 * max(1.0, float(back_ref.abs().max()))


@pytest.mark.parametrize("mode", ["pca", "chol", "sym", "cdf"])
def test_ot_step_on_real_conv4_features(ob, ref, mode):
    """conv4_1 features of the bundled images (real weights), PCA-projected like optex.py:65-67,110: the
    reference's optimal_transport on CPU vs the B200 step, same injected rotation."""
    style = load_style(ref, "graffiti.jpg", 256)
    other = load_style(ref, "zebra.jpg", 256)
    with reference.in_reference_dir(), torch.inference_mode():
        enc = ref.vgg.Encoder(4)
        fs, fp = enc(style), enc(other)
        fs, v = ref.optex.fit_pca(fs)
        fp = fp @ v
        c = fs.shape[-1]
        rot = torch.from_numpy(__import__("oracle.rotation", fromlist=["x"]).haar_rotation_qr(c, 3))
        saved = ref.optex.random_rotation
        ref.optex.random_rotation = lambda N, device="cpu", impl="scipy": rot          # inject (optex.py:168)
        try:
            want = ref.optex.optimal_transport(fp, fs, mode)
        finally:
            ref.optex.random_rotation = saved
    got = ob.optimal_transport(fp.cuda(), fs.cuda(), mode, rotation=rot.float().cuda()).cpu()
    scale = max(1.0, float(want.abs().max()))
    d = (got - want).abs()
    if mode == "cdf":     # discontinuous map: bulk criterion (tests/test_gpu_baseline_shapes.py bounds the outliers)
        assert float((d > 2e-4 * scale).float().mean()) <= 5e-3
    else:
        assert float(d.max()) <= 5e-4 * scale, f"{mode}: {float(d.max())} vs scale {scale}"


def test_reference_forward_is_the_caller_after_install(ob, ref):
    """INTEGRATION.md's drop-in: the reference's OptimalTexture.forward, unmodified, with its module-level names
    rebound by ob.install().  One 256^2 pass, hist pca (a continuous map, and rotation-invariant - so no rotation
    stream has to be injected), real weights, bundled style image."""
    optex = ref.optex
    torch.manual_seed(0)
    style = load_style(ref, "graffiti.jpg", 256)
    pastiche = torch.rand(1, 3, 256, 256)
    kw = dict(size=256, iters=60, passes=1, hist_mode="pca")
    with reference.in_reference_dir(), torch.inference_mode():
        model = optex.OptimalTexture(**kw)
        want = model.forward(pastiche, [style])
    names = ["optimal_transport", "hist_match", "random_rotation", "fit_pca", "mix_style_features", "resize",
             "rgb_to_hls", "hls_to_rgb"]
    saved = {n: getattr(optex, n) for n in names}
    from optimaltextures_b200 import _lib

    l0 = _lib.lib().optex_launch_count()
    try:
        ob.install(optex)
        assert optex.optimal_transport is ob.optimal_transport
        with torch.inference_mode():
            got = model.to("cuda").forward(pastiche.cuda(), [style.cuda()]).cpu()
    finally:
        for n, f in saved.items():
            setattr(optex, n, f)
        model.to("cpu")
    assert _lib.lib().optex_launch_count() - l0 > 100, "the B200 library did not run under the reference's forward"
    assert got.shape == want.shape and bool(torch.isfinite(got).all())
    d = (got - want).abs()
    # 5 layers x (encode, PCA, ~12 OT steps, decode) chained: fp32 differences grow through the decoders
    assert float(d.mean()) <= 2e-3 and float(d.max()) <= 5e-2, (float(d.mean()), float(d.max()))
    assert abs(float(got.mean()) - float(want.mean())) <= 1e-3
    assert abs(float(got.std()) - float(want.std())) <= 1e-3
