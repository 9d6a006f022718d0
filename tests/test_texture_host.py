"""Host logic of the synthesis loop (no GPU): the util mirror against values produced by the REAL reference
(tests/golden/util.json, misc.npz), and the explicit resize / HLS arithmetic of oracle/image_oracle.py."""
import json
import os
from argparse import Namespace

import numpy as np
import pytest
import torch

from optimaltextures_b200 import util as outil
from oracle import image_oracle, texture_oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _util_golden():
    with open(os.path.join(HERE, "golden", "util.json")) as fh:
        return json.load(fh)


def test_get_size_matches_reference():
    for args, want in _util_golden()["get_size"]:
        assert list(outil.get_size(*args)) == want
        assert list(texture_oracle.get_size(*args)) == want


def test_output_names_match_reference():
    for case in _util_golden()["names"]:
        ns = Namespace(**case["args"])
        stem = f"{ns.output_dir}/{outil.output_name(ns)}"
        want = case["paths"]
        got = [stem + (f"_{o + 1}" if case["batch"] > 1 else "") + ".png" for o in range(case["batch"])]
        assert got == want


def test_schedule_matches_reference(golden):
    g = golden("misc")
    for (size, iters, passes) in ((512, 500, 5), (256, 500, 4), (1024, 500, 5), (2048, 300, 3)):
        its, sizes = outil.get_iters_and_sizes(size, iters, passes, True)
        np.testing.assert_array_equal(np.asarray(its), g[f"sched_{size}_{iters}_{passes}_iters"])
        np.testing.assert_array_equal(np.asarray(sizes), g[f"sched_{size}_{iters}_{passes}_sizes"])
    # effective per-layer iterations of `--iters 500 --passes 5` with the [l - 1] quirk
    its, _ = outil.get_iters_and_sizes(512, 500, 5, True)
    assert [its[0][l - 1] for l in range(5)] == [40, 8, 13, 22, 40]
    assert sum(its[p][l - 1] for p in range(5) for l in range(5)) == 493


def test_no_multires_schedule_is_usable():
    """The reference raises here (util.py:86, list.tolist()); the mirror returns the documented intent."""
    its, sizes = outil.get_iters_and_sizes(512, 500, 5, False)
    assert sizes == [512] * 5 and len(its) == 5 and all(len(r) == 5 for r in its)
    assert its[0] == [int(100 * p) for p in (np.array([128, 192, 320, 576, 576]) / 1792)]


def test_round32_and_layout_helpers():
    assert [outil.round32(v) for v in (1, 32, 33, 255, 256, 257)] == [32, 32, 64, 256, 256, 288]
    x = torch.zeros(2, 3, 5, 7)
    assert outil.to_nhwc(x).shape == (2, 5, 7, 3) and outil.to_nchw(outil.to_nhwc(x)).shape == x.shape
    assert outil.name("a/b/zebra.small.jpg") == "zebra"


def test_resize_requires_gpu():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        outil.resize(torch.zeros(1, 3, 8, 8), (4, 4))


@pytest.mark.parametrize("shape", [(64, 48, 32, 32), (37, 53, 64, 96), (256, 256, 448, 448), (416, 416, 256, 224),
                                   (100, 80, 33, 95), (32, 32, 32, 64), (5, 7, 1, 1)])
def test_resize_explicit_matches_torch(shape):
    """The written-out filter (what csrc/image.cu implements) == torch's CPU interpolate, to fp32 rounding."""
    h, w, ho, wo = shape
    x = torch.rand(2, 3, h, w, generator=torch.Generator().manual_seed(h * w))
    want = image_oracle.resize(x, (ho, wo)).numpy()
    got = image_oracle.resize_explicit(x.numpy(), (ho, wo))
    np.testing.assert_allclose(got, want, rtol=0, atol=4e-6)


def test_hls_round_trip_and_known_colours():
    rgb = torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0.5, 0.5, 0.5], [0, 0, 0], [1, 1, 1],
                        [0.2, 0.4, 0.6]]).T.reshape(1, 3, 1, 7)
    hls = image_oracle.rgb_to_hls(rgb)
    np.testing.assert_allclose(hls[0, 0, 0, :3].numpy(), [0, 2 * np.pi / 3, 4 * np.pi / 3], atol=1e-6)   # hue
    np.testing.assert_allclose(hls[0, 1, 0].numpy(), [0.5, 0.5, 0.5, 0.5, 0, 1, 0.4], atol=1e-6)         # lightness
    np.testing.assert_allclose(hls[0, 2, 0].numpy(), [1, 1, 1, 0, 0, 0, 0.5], atol=1e-6)                 # saturation
    x = torch.rand(2, 3, 16, 16, generator=torch.Generator().manual_seed(3))
    np.testing.assert_allclose(image_oracle.hls_to_rgb(image_oracle.rgb_to_hls(x)).numpy(), x.numpy(), atol=2e-6)
    same = image_oracle.lightness_transfer(x, x)
    np.testing.assert_allclose(same.numpy(), x.numpy(), atol=2e-6)


# ---------------------------------------------------------------------------------------------- control flow
@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("name", ["synth_pca", "mix_content_chol_opt", "nopca_cdf_lum"])
def test_product_control_flow_matches_reference(golden, monkeypatch, name, overlap):
    """The product's `OptimalTexture` (optimaltextures_b200/texture.py) with each of its KERNEL calls replaced by the
    oracle's op for that stage reproduces the real reference's output (tests/golden/texture.npz) bit for bit: the
    orchestration - pass sizes, resize rule, per-layer iteration counts with the [l - 1] quirk, content strengths
    and the l <= 2 rule, mixing mask, colour-transfer branches, the order of RNG / rotation consumption - is the
    reference's.  (The kernels themselves are compared stage by stage on the GPU, tests/test_gpu_texture.py.)
    overlap=True: the host-side order of the `overlap_style` schedule (pass p + 1's style side launched on a side
    stream before pass p's layers, finished after them) consumes the RNG and produces every tensor identically."""
    from optimaltextures_b200 import texture
    from oracle import ot_oracle, texture_cases, vgg_oracle

    class Enc:
        def __init__(self, d, state_dict=None, models_dir=None, device=None):
            self.depth, self.sd, self.layers = d, state_dict, []

        def forward_all(self, x):
            return vgg_oracle.encoder_forward(x, self.sd, self.depth, all_depths=True)

        def __call__(self, x):
            return vgg_oracle.encoder_forward(x, self.sd, self.depth)

    class Dec:
        def __init__(self, d, state_dict=None, models_dir=None, device=None):
            self.depth, self.sd = d, state_dict

        def __call__(self, x):
            return vgg_oracle.decoder_forward(x, self.sd, self.depth)

    def ot_loop(f, s, mode, iters, rotations=None, content=None, content_strength=0.0):
        for r in rotations:
            f = ot_oracle.ot_step(f, s, r, mode)
            if content is not None:
                f = f + content_strength * (content - f)
        return f

    monkeypatch.setattr(texture, "require_cuda", lambda *t: torch.device("cpu"))
    monkeypatch.setattr(texture._util, "resize", image_oracle.resize)
    monkeypatch.setattr(texture._vgg, "Encoder", Enc)
    monkeypatch.setattr(texture._vgg, "Decoder", Dec)
    monkeypatch.setattr(texture._optex, "pca_project", lambda x, v, transpose=False: x @ (v.T if transpose else v))
    monkeypatch.setattr(texture._optex, "ot_loop", ot_loop)
    monkeypatch.setattr(texture, "recentre", image_oracle.recentre)
    monkeypatch.setattr(texture, "lightness_transfer", image_oracle.lightness_transfer)
    monkeypatch.setattr(texture, "mix_style_features", texture_oracle.mix_style_features)

    events = []
    if overlap:
        # drive the stream branch of the overlapped schedule too, with stand-ins that log the synchronisation calls
        import contextlib

        class FakeStream:
            def __init__(self, tag):
                self.tag = tag

            def wait_event(self, ev):
                events.append((self.tag, "wait", ev))

            def record_event(self):
                events.append((self.tag, "record", len(events)))
                return f"{self.tag}-event-{len(events)}"

        class FakeLib:
            slot = 0

            def optex_set_scratch_slot(self, slot):
                prev, FakeLib.slot = FakeLib.slot, slot
                events.append(("lib", "slot", slot))
                return prev

        main, side = FakeStream("main"), FakeStream("side")
        monkeypatch.setattr(texture.OptimalTexture, "_streams_for", lambda self, p: (main, side, main.record_event()))
        monkeypatch.setattr(texture.torch.cuda, "stream", lambda s: contextlib.nullcontext())
        monkeypatch.setattr(texture._lib, "lib", lambda: FakeLib())

    g = golden("texture")
    kwargs, styles, content, pastiche = texture_cases.texture_inputs(name)
    model = texture.OptimalTexture(state_dicts=texture_cases.state_dicts(), rotations=texture_cases.texture_rotation,
                                   device="cpu", pca=ot_oracle.fit_pca, overlap_style=overlap, **kwargs)
    model.pad_channels = 1          # the zero-channel padding is exact in real arithmetic, not bit for bit
    torch.manual_seed(77)
    with torch.inference_mode():
        out = model.forward(pastiche, styles, content)
    assert model.ot_calls == int(g[f"{name}_calls"])
    np.testing.assert_array_equal(out.contiguous().numpy(), g[f"{name}_out"])
    if overlap:
        passes = kwargs["passes"]
        # per pass: launch (slot 1, side waits for the start event, slot 0) and finish (slot 1, side records "ready",
        # slot 0); main waits for the ready event; the launch of pass p + 1 sits between pass p's finish and its layers
        assert [e for e in events if e[0] == "lib"] == [("lib", "slot", 1), ("lib", "slot", 0)] * (2 * passes)
        assert sum(1 for e in events if e[:2] == ("side", "wait")) == passes
        ready = [f"side-event-{i + 1}" for i, e in enumerate(events) if e[:2] == ("side", "record")]
        assert [e[2] for e in events if e[:2] == ("main", "wait")] == ready and len(ready) == passes


def test_cli_flags_match_the_reference():
    """Every flag of optex.py:223-243, with the reference's defaults (values read off the reference's parser)."""
    from optimaltextures_b200.__main__ import build_parser

    a = build_parser().parse_args([])
    want = dict(style=["style/graffiti.jpg"], content=None, batch=1, size=512, passes=5, iters=500, hist_mode="chol",
                color_transfer=None, content_strength=0.01, style_scale=1.0, mixing_alpha=0.5, no_pca=False,
                no_multires=False, seed=None, no_tf32=False, cudnn_benchmark=False, compile=False, script=False,
                device=None, memory_format="contiguous", output_dir="output/")
    for k, v in want.items():
        assert getattr(a, k) == v, k
    b = build_parser().parse_args(["-s", "a.jpg", "b.jpg", "-c", "c.jpg", "--hist_mode", "cdf", "--color_transfer", "opt"])
    assert b.style == ["a.jpg", "b.jpg"] and b.content == "c.jpg" and b.hist_mode == "cdf"
    with pytest.raises(SystemExit):
        build_parser().parse_args(["--hist_mode", "sort"])          # the CLI exposes the reference's four modes


def test_load_and_save_image_round_trip(tmp_path):
    """util.load_image (PIL decode + Lanczos resize, util.py:27-30) and save_image's PNG naming, on the host."""
    from PIL import Image

    arr = (np.random.RandomState(0).rand(70, 90, 3) * 255).astype(np.uint8)
    path = tmp_path / "zebra.png"
    Image.fromarray(arr).save(path)
    x = outil.load_image(str(path), 64, device="cpu")
    assert x.shape == (1, 3, *reversed(outil.get_size(64, 1, 90, 70, True))) or x.dim() == 4
    assert 0.0 <= float(x.min()) and float(x.max()) <= 1.0
    ns = Namespace(style=[str(path)], content=None, mixing_alpha=0.5, content_strength=0.01, hist_mode="chol",
                   no_pca=False, no_multires=False, style_scale=1.0, color_transfer=None, size=64,
                   output_dir=str(tmp_path))
    (out,) = outil.save_image(x, ns)
    assert out.endswith("zebra_cholhist_64.png")
    back = np.asarray(Image.open(out))
    assert back.shape[2] == 3 and np.abs(back.astype(int) - (x[0].permute(1, 2, 0).numpy() * 255 + 0.5).astype(int)).max() <= 1


def test_fit_pca_basis_argument_checks():
    """host-side validation of the additive basis= / warm= arguments of fit_pca (no device call)."""
    import pytest
    import torch

    from optimaltextures_b200 import optex as gpu

    cpu = torch.device("cpu")
    gpu._check_basis(None, 8, cpu, False)
    gpu._check_basis(torch.zeros(8, 8, dtype=torch.float64), 8, cpu, True)
    with pytest.raises(ValueError):
        gpu._check_basis(None, 8, cpu, True)                                   # warm start without a basis
    with pytest.raises(ValueError):
        gpu._check_basis(torch.zeros(8, 8), 8, cpu, False)                     # fp32
    with pytest.raises(ValueError):
        gpu._check_basis(torch.zeros(8, 4, dtype=torch.float64), 8, cpu, False)
    with pytest.raises(ValueError):
        gpu._check_basis(torch.zeros(8, 16, dtype=torch.float64)[:, ::2], 8, cpu, False)   # not contiguous


def test_pca_warm_start_bookkeeping(monkeypatch):
    """OptimalTexture hands each layer's float64 basis buffer to fit_pca_many: cold in pass 0 of every forward(), warm in
    the later passes, one [c, c] buffer per layer kept across passes (device PCA stubbed by the oracle's fit_pca)."""
    from optimaltextures_b200 import texture
    from oracle import ot_oracle, texture_cases, vgg_oracle

    class Enc:
        def __init__(self, d, state_dict=None, models_dir=None, device=None):
            self.depth, self.sd, self.layers = d, state_dict, []

        def forward_all(self, x):
            return vgg_oracle.encoder_forward(x, self.sd, self.depth, all_depths=True)

        def __call__(self, x):
            return vgg_oracle.encoder_forward(x, self.sd, self.depth)

    class Dec:
        def __init__(self, d, state_dict=None, models_dir=None, device=None):
            self.depth, self.sd = d, state_dict

        def __call__(self, x):
            return vgg_oracle.decoder_forward(x, self.sd, self.depth)

    calls = []

    def launch(tensors, *, bases=None, warm=None):
        calls.append((list(warm), [None if b is None else b.data_ptr() for b in bases],
                      [None if b is None else tuple(b.shape) for b in bases], None if bases[0] is None else bases[0].dtype))
        return tensors

    def finish(state, *, round_k_to=1, sweeps_out=None):
        sweeps_out.extend([7] * len(state))
        return [ot_oracle.fit_pca(t) for t in state]

    monkeypatch.setattr(texture, "require_cuda", lambda *t: torch.device("cpu"))
    monkeypatch.setattr(texture._util, "resize", image_oracle.resize)
    monkeypatch.setattr(texture._vgg, "Encoder", Enc)
    monkeypatch.setattr(texture._vgg, "Decoder", Dec)
    monkeypatch.setattr(texture._optex, "pca_project", lambda x, v, transpose=False: x @ (v.T if transpose else v))
    monkeypatch.setattr(texture._optex, "ot_loop", lambda f, *a, **k: f)
    monkeypatch.setattr(texture._optex, "fit_pca_many_launch", launch)
    monkeypatch.setattr(texture._optex, "fit_pca_many_finish", finish)
    kwargs, styles, content, pastiche = texture_cases.texture_inputs("synth_pca")
    model = texture.OptimalTexture(state_dicts=texture_cases.state_dicts(), device="cpu", pca_warm_start=True,
                                   overlap_style=False, **kwargs)
    model.pad_channels = 1
    with torch.inference_mode():
        model.forward(pastiche, styles, content)
        assert [c[0] for c in calls] == [[False] * 5, [True] * 5]
        assert calls[0][1] == calls[1][1], "the basis buffers must persist from pass to pass"
        assert calls[0][2] == [(512, 512), (512, 512), (256, 256), (128, 128), (64, 64)] and calls[0][3] == torch.float64
        assert model.pca_sweeps == [[7] * 5, [7] * 5]
        model.forward(pastiche, styles, content)
    assert [c[0] for c in calls[2:]] == [[False] * 5, [True] * 5], "a new forward() starts cold"
    # the default: cold solves without a basis (the blocked-order solver of pca.cu)
    cold = texture.OptimalTexture(state_dicts=texture_cases.state_dicts(), device="cpu", overlap_style=False, **kwargs)
    cold.pad_channels = 1
    del calls[:]
    with torch.inference_mode():
        cold.forward(pastiche, styles, content)
    assert [c[0] for c in calls] == [[False] * 5, [False] * 5] and calls[0][1] == [None] * 5


def test_equal_size_passes_share_the_style_side(monkeypatch):
    """Passes of equal size refit the same style at the same size (optex.py:62-67 is deterministic): the product
    encodes / solves once per distinct size within a forward() and reuses the result."""
    from optimaltextures_b200 import texture
    from oracle import ot_oracle, texture_cases, vgg_oracle

    class Enc:
        def __init__(self, d, state_dict=None, models_dir=None, device=None):
            self.depth, self.sd, self.layers = d, state_dict, []

        def forward_all(self, x):
            return vgg_oracle.encoder_forward(x, self.sd, self.depth, all_depths=True)

        def __call__(self, x):
            return vgg_oracle.encoder_forward(x, self.sd, self.depth)

    class Dec:
        def __init__(self, d, state_dict=None, models_dir=None, device=None):
            self.depth, self.sd = d, state_dict

        def __call__(self, x):
            return vgg_oracle.decoder_forward(x, self.sd, self.depth)

    fits = []

    def pca(t):
        fits.append(tuple(t.shape))
        return ot_oracle.fit_pca(t)

    monkeypatch.setattr(texture, "require_cuda", lambda *t: torch.device("cpu"))
    monkeypatch.setattr(texture._util, "resize", image_oracle.resize)
    monkeypatch.setattr(texture._vgg, "Encoder", Enc)
    monkeypatch.setattr(texture._vgg, "Decoder", Dec)
    monkeypatch.setattr(texture._optex, "pca_project", lambda x, v, transpose=False: x @ (v.T if transpose else v))
    monkeypatch.setattr(texture._optex, "ot_loop", lambda f, *a, **k: f)
    kwargs, styles, content, pastiche = texture_cases.texture_inputs("synth_pca")
    kwargs = dict(kwargs, size=64, passes=3)          # sizes = linspace(256, 64, 3) rounded: all distinct ...
    model = texture.OptimalTexture(state_dicts=texture_cases.state_dicts(), device="cpu", pca=pca, overlap_style=False,
                                   **kwargs)
    model.sizes = [64, 64, 96]                        # ... so force two equal passes and one different
    model.pad_channels = 1
    with torch.inference_mode():
        model.forward(pastiche, styles, content)
    assert len(fits) == 2 * 5, fits                   # five layers for size 64 (once), five for size 96
