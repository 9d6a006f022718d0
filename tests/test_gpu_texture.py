"""GPU parity: the synthesis loop (`OptimalTexture.forward`, optex.py:81-139) on the B200 kernels vs the CPU oracle
(oracle/texture_oracle.py, pinned BIT-EXACTLY to the real reference's forward on these very cases in
tests/test_oracle_golden.py::test_texture_forward_matches_reference).

How the loop is compared.  With random-weight networks the loop as a whole is chaotic: the ORACLE run twice, with
its input pastiche perturbed by 1e-6 (relative), ends 1e-3 (mean) / 1e-2 (max) apart, a 1e-4 perturbation 3e-2 / 0.3
(measured, fp32 CPU) - no two fp32 implementations of it, the reference on two BLAS builds included, agree
element-wise at the end.  So parity is stated where it is well posed:
  1. control flow: the product's `OptimalTexture` with its kernels replaced by the oracle's ops reproduces the real
     reference's output bit for bit (tests/test_texture_host.py, CPU);
  2. every stage, TEACHER-FORCED: the oracle records each stage's inputs and output (`trace`), the B200 stage is fed
     the oracle's inputs and compared with the oracle's output - this file;
  3. end to end: finite, deterministic, the reference's number of OT calls, and output statistics close to the
     oracle's (a sanity bound, not a parity claim).

Stated tolerances (floating point), relative to max |reference stage output| unless noted:
  resize 1e-5 (absolute, inputs in [0,1]) | encode / decode 1e-3 (13 conv layers of 3xTF32 GEMMs, see test_gpu_vgg.py)
  project / unproject 2e-5 | recentre 2e-6 | mix (two chol / cdf hist_match + blend) 1e-3, cdf in bulk
  fit_pca: k equal, projector within 2e-4 (see test_gpu_pca.py)
  ot_step / ot_loop, smooth modes (pca, chol): 2e-3 - the oracle's own fp32 answer is 7e-4 from its fp64 answer on
    the worst-conditioned of these steps (covariance eigenvalues 1 .. 2.6e5)
  ot_step cdf: >= 99 % of the elements within 2e-3 * (99.9th percentile of |reference|): the map is discontinuous
    at bin edges, and the reference's `interp` (histmatch.py:72-92) extrapolates isolated elements to 1e5 x the
    data range, so neither max-norm nor max-scale is meaningful; multi-iteration cdf loops are compared step by step
  lightness transfer 2e-5 * scale.
The measured errors are written to gpurun_out/texture_parity.json."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import texture_cases, texture_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["synth_pca", "mix_content_chol_opt", "nopca_cdf_lum"]


@pytest.fixture(scope="module")
def ob():
    import optimaltextures_b200 as ob

    return ob


def _record(name, stats):
    path = os.path.join(ROOT, "gpurun_out", "texture_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = stats
        json.dump(data, open(path, "w"), indent=1)
    except OSError:
        pass


def _mask_noise():
    noise = {}

    def fn(shape):
        if shape not in noise:
            noise[shape] = torch.rand(shape, generator=torch.Generator().manual_seed(99))
        return noise[shape]

    return fn


def _oracle_trace(name):
    kwargs, styles, content, pastiche = texture_cases.texture_inputs(name)
    model = texture_oracle.OptimalTexture(texture_cases.state_dicts(), rotation_fn=texture_cases.texture_rotation,
                                          mask_fn=_mask_noise(), **kwargs)
    model.trace = []
    with torch.inference_mode():
        out = model.forward(pastiche.clone(), [s.clone() for s in styles], None if content is None else content.clone())
    return model, out


def rel(got, want):
    got, want = got.double().cpu(), want.double()
    assert got.shape == want.shape, f"{tuple(got.shape)} vs {tuple(want.shape)}"
    return float((got - want).abs().max() / max(1.0, float(want.abs().max())))


def bulk(got, want, tol=2e-3):
    got, want = got.double().cpu().flatten(), want.double().flatten()
    scale = max(1.0, float(torch.quantile(want.abs()[:: max(1, want.numel() // 1000000)], 0.999)))
    return float(((got - want).abs() <= tol * scale).double().mean())


@pytest.mark.parametrize("name", CASES)
def test_every_stage_teacher_forced(ob, name):
    from optimaltextures_b200 import texture, util, vgg

    model, _ = _oracle_trace(name)
    sd = texture_cases.state_dicts()
    enc = {d: vgg.Encoder(d, state_dict=sd[("encoder", d)]) for d in range(1, 6)}
    dec = {d: vgg.Decoder(d, state_dict=sd[("decoder", d)]) for d in range(1, 6)}
    worst = {}

    def note(stage, value, limit, lower_is_better=True):
        key = stage
        worst[key] = max(worst.get(key, 0.0), value) if lower_is_better else min(worst.get(key, 1.0), value)
        ok = value <= limit if lower_is_better else value >= limit
        assert ok, f"{name}: stage {stage}: {value:.3e} vs limit {limit:.3e}"

    seen = set()
    for entry in model.trace:
        stage = entry[0]
        seen.add(stage)
        if stage == "resize":
            _, x, size, out = entry
            note("resize", float((util.resize(x.cuda(), size).cpu() - out).abs().max()), 1e-5)
        elif stage == "encode":
            _, d, x, out = entry
            note("encode", rel(enc[d](x.cuda()), out), 1e-3)
        elif stage == "decode":
            _, d, f, out = entry
            note("decode", rel(dec[d](f.cuda()), out), 1e-3)
        elif stage == "fit_pca":
            _, raw, eig, feats = entry
            f_gpu, e_gpu = ob.fit_pca(raw.cuda())
            assert e_gpu.shape == eig.shape, f"{name}: fit_pca kept {e_gpu.shape[1]} components, reference {eig.shape[1]}"
            e_gpu = e_gpu.cpu()
            note("fit_pca_projector", float((e_gpu @ e_gpu.T - eig @ eig.T).abs().max()), 2e-4)
        elif stage == "project":
            _, x, eig, transpose, out = entry
            note("project", rel(ob.pca_project(x.cuda(), eig.cuda(), transpose=transpose), out), 2e-5)
        elif stage == "recentre":
            _, cf, sf, out = entry
            note("recentre", rel(texture.recentre(cf.cuda(), sf.cuda()), out), 2e-6)
        elif stage == "mix":
            _, sf, mask, alpha, mode, out = entry
            got = texture.mix_style_features([sf.cuda()], mask.cuda(), alpha, mode)[0]
            if mode == "cdf":
                note("mix_bulk", bulk(got, out), 0.99, lower_is_better=False)
            else:
                note("mix", rel(got, out), 1e-3)
        elif stage == "ot_step":
            _, f, style, rot, mode, out = entry
            got = ob.optimal_transport(f.cuda(), style.cuda(), mode, rotation=rot.float().cuda())
            if mode == "cdf":
                note(f"ot_step_cdf_bulk(c={f.shape[-1]})", bulk(got, out), 0.99, lower_is_better=False)
            else:
                note("ot_step", rel(got, out), 2e-3)
        elif stage == "ot_loop":
            _, f, style, first, n, mode, content, strength, out = entry
            if mode == "cdf":
                continue                      # compared step by step above
            rots = torch.stack([texture_cases.texture_rotation(f.shape[-1], first + i).float() for i in range(n)])
            got = ob.ot_loop(f.cuda(), style.cuda(), mode, n, rotations=rots.cuda(),
                             content=None if content is None else content.cuda(), content_strength=strength)
            note("ot_loop", rel(got, out), 2e-3)
        elif stage == "lightness":
            _, content, pastiche, out = entry
            note("lightness", rel(texture.lightness_transfer(content.cuda(), pastiche.cuda()), out), 2e-5)
        else:
            raise AssertionError(f"unknown stage {stage}")
    _record(f"stages_{name}", worst)
    assert {"encode", "decode", "ot_step"} <= seen


def _run_product(ob, name, **override):
    from optimaltextures_b200 import texture

    kwargs, styles, content, pastiche = texture_cases.texture_inputs(name)
    kwargs = dict(kwargs, **override)
    model = texture.OptimalTexture(state_dicts=texture_cases.state_dicts(), rotations=texture_cases.texture_rotation,
                                   mixing_noise=_mask_noise(), **kwargs)
    assert model.shared_encoder
    out = model.forward(pastiche.cuda(), [s.cuda() for s in styles], None if content is None else content.cuda())
    torch.cuda.synchronize()
    return model, out


@pytest.mark.parametrize("name", ["synth_pca", "mix_content_chol_opt"])
def test_end_to_end_sanity(ob, name):
    """Same injections on both sides; the chaotic loop is compared through what the algorithm controls: the output's
    per-channel statistics (and the bookkeeping: shapes, call counts, determinism)."""
    ref_model, want = _oracle_trace(name)
    model, got = _run_product(ob, name)
    model2, got2 = _run_product(ob, name)
    assert model.ot_calls == ref_model.ot_calls
    assert got.shape == want.shape and torch.isfinite(got).all()
    assert torch.equal(got, got2), "the loop is not deterministic"
    got = got.cpu().double()
    want = want.double()
    scale = float(want.abs().max())
    stats = {"mean_abs_err_over_scale": float((got - want).abs().mean()) / scale,
             "channel_mean_err": float((got.mean((0, 2, 3)) - want.mean((0, 2, 3))).abs().max()) / scale,
             "channel_std_ratio": (got.std((0, 2, 3)) / want.std((0, 2, 3))).tolist()}
    _record(f"e2e_{name}", stats)
    # measured: 0.024 / 0.11 mean |err| (the oracle against ITSELF with a 1e-4 input perturbation: 0.034 / 0.066)
    assert stats["mean_abs_err_over_scale"] <= 0.25
    assert stats["channel_mean_err"] <= 0.05
    assert all(0.75 <= r <= 1.33 for r in stats["channel_std_ratio"])


def test_device_rotations_and_schedule(ob):
    """Without injection the loop draws its rotations on the device: runs, is deterministic under manual_seed, and
    consumes exactly the reference's number of OT calls."""
    from optimaltextures_b200 import texture

    kwargs, styles, content, pastiche = texture_cases.texture_inputs("synth_pca")
    sd = texture_cases.state_dicts()
    outs = []
    for _ in range(2):
        ob.manual_seed(11)
        model = texture.OptimalTexture(state_dicts=sd, **kwargs)
        outs.append(model.forward(pastiche.cuda(), [s.cuda() for s in styles]))
        its = model.iters_per_pass_and_layer
        assert model.ot_calls == sum(its[p][l - 1] for p in range(model.passes) for l in range(5))
    assert outs[0].shape == (1, 3, 64, 64) and torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("name", ["nopca_cdf_lum", "mix_content_chol_opt"])
def test_colour_transfer_and_content_paths_run(ob, name):
    model, got = _run_product(ob, name)
    kwargs, styles, content, pastiche = texture_cases.texture_inputs(name)
    assert got.shape == content.shape and torch.isfinite(got).all()


def test_no_multires_runs(ob):
    from optimaltextures_b200 import texture

    sd = texture_cases.state_dicts()
    model = texture.OptimalTexture(size=64, iters=10, passes=2, hist_mode="chol", no_multires=True, state_dicts=sd)
    assert model.sizes == [64, 64]
    out = model.forward(torch.rand(1, 3, 64, 64, device="cuda"), [torch.rand(1, 3, 64, 96, device="cuda")])
    assert out.shape == (1, 3, 64, 64) and torch.isfinite(out).all()


def test_cpu_tensors_are_refused(ob):
    from optimaltextures_b200 import texture

    model = texture.OptimalTexture(size=32, iters=5, passes=1, state_dicts=texture_cases.state_dicts())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.forward(torch.rand(1, 3, 32, 32), [torch.rand(1, 3, 32, 32)])


@pytest.mark.parametrize("mode", ["pca", "cdf"])
def test_overlapped_schedule_is_bit_identical(ob, mode):
    """The default schedule (the style side of pass p + 1 - encoders and the five cooperative PCA solves - on a side
    stream beside pass p's OT loops) gives the image of the serial schedule bit for bit, run after run; passes of
    equal size share one style-side result (one set of Jacobi solves per distinct size)."""
    from optimaltextures_b200 import texture

    sd = texture_cases.state_dicts()
    g = torch.Generator().manual_seed(3)
    style = torch.rand(1, 3, 96, 64, generator=g).cuda()
    pastiche = torch.rand(1, 3, 64, 64, generator=g).cuda()
    outs = {}
    for overlap in (False, True, True):
        model = texture.OptimalTexture(size=128, iters=60, passes=3, hist_mode=mode, state_dicts=sd,
                                       overlap_style=overlap)
        ob.manual_seed(5)
        out = model.forward(pastiche, [style])
        assert bool(torch.isfinite(out).all())
        if overlap in outs:
            assert torch.equal(out, outs[overlap]), "two runs of the overlapped schedule differ"
        outs[overlap] = out
        assert len(model.pca_sweeps) == 3                      # sizes 256, 192, 128: three distinct style sides
    assert torch.equal(outs[False], outs[True])
    same = texture.OptimalTexture(size=64, iters=40, passes=3, hist_mode=mode, no_multires=True, state_dicts=sd)
    ob.manual_seed(5)
    out = same.forward(pastiche, [style])
    assert bool(torch.isfinite(out).all()) and len(same.pca_sweeps) == 1   # three passes of 64^2: one set of solves


def test_fit_pca_many_equals_fit_pca(ob):
    """The concurrent per-layer PCA of a pass (side streams) == five sequential `fit_pca` calls, bit for bit."""
    g = torch.Generator().manual_seed(4)
    blocks = [torch.relu(torch.randn(1, h, h, c, generator=g) * torch.linspace(0.2, 2.0, c)).cuda()
              for h, c in ((8, 512), (16, 512), (32, 256), (64, 128), (128, 64))]
    from optimaltextures_b200 import optex

    many = optex.fit_pca_many(blocks)
    for x, (f, e) in zip(blocks, many):
        f1, e1 = ob.fit_pca(x)
        assert e.shape == e1.shape and torch.equal(e, e1) and torch.equal(f, f1)


@pytest.mark.parametrize("mode", ["pca", "sym", "chol"])
def test_zero_channel_padding_is_exact(ob, mode):
    """OptimalTexture pads the loops with zero channels up to a multiple of 32 (tensor-core path) and rotates by
    diag(R_c, I): same result as the unpadded loop to fp32 rounding (covariances become blockdiag(Sigma + I, I))."""
    from optimaltextures_b200 import texture

    g = torch.Generator().manual_seed(9)
    c = 49
    f = torch.relu(torch.randn(1, 40, 40, c, generator=g)).cuda()
    s = torch.relu(1.5 * torch.randn(1, 36, 44, c, generator=g) + 0.25).cuda()
    content = torch.relu(torch.randn(1, 40, 40, c, generator=g)).cuda()
    model = texture.OptimalTexture(size=32, iters=5, passes=1, state_dicts=texture_cases.state_dicts(),
                                   rotations=texture_cases.texture_rotation)
    outs = {}
    for pad in (32, 1):
        model.pad_channels = pad
        model.ot_calls = 0
        outs[pad] = model._ot_layer(f, s, mode, 4, content, 0.05)
    assert outs[32].shape == outs[1].shape == f.shape
    err = float((outs[32] - outs[1]).abs().max() / outs[1].abs().max())
    assert err <= (1e-4 if mode != "chol" else 5e-4), err


@pytest.mark.parametrize("mode", ["cdf", "sort"])
@pytest.mark.parametrize("c", [23, 85, 181])
def test_zero_channel_padding_per_channel_modes(ob, mode, c):
    """cdf / sort at PCA'd channel counts: ONE step through the padded path (diag(R_c, I), tensor-core GEMMs) against
    the unpadded step (SIMT GEMMs).  The padded channels never mix with the real ones; the two differ only by the GEMM
    arithmetic (3xTF32 vs fp32 FFMA), i.e. like any two fp32-grade rotations: bulk criterion of the OT step."""
    from optimaltextures_b200 import texture

    g = torch.Generator().manual_seed(c)
    f = (torch.randn(1, 64, 48, c, generator=g) * torch.logspace(1, -0.5, c) + 0.3).cuda()
    s = ((1.2 * torch.randn(1, 50, 60, c, generator=g) + 0.1) * torch.logspace(1, -0.5, c) + 0.3).cuda()
    model = texture.OptimalTexture(size=32, iters=5, passes=1, state_dicts=texture_cases.state_dicts(),
                                   rotations=texture_cases.texture_rotation)
    outs = {}
    for pad in (32, 1):
        model.pad_channels = pad
        model.ot_calls = 0
        outs[pad] = model._ot_layer(f, s, mode, 1, None, 0.0)
    # compared in the rotated frame, where a bin / rank flip touches ONE element (in the output frame the inverse
    # rotation spreads it over the pixel's c channels)
    r = texture_cases.texture_rotation(c, 0).double().cuda()
    ma, mb = outs[32].reshape(-1, c).double() @ r, outs[1].reshape(-1, c).double() @ r
    d = (ma - mb).abs()
    scale = float(mb.abs().max())
    assert float((d > 2e-4 * scale).double().mean()) <= 5e-3, float((d > 2e-4 * scale).double().mean())
    # device-drawn rotations: the padded loop draws c x c matrices from the same (seed, counter) stream
    model.rotations = None
    ob.manual_seed(4)
    a = model._ot_layer(f, s, mode, 3, None, 0.0)
    ob.manual_seed(4)
    b = model._ot_layer(f, s, mode, 3, None, 0.0)
    assert torch.equal(a, b) and bool(torch.isfinite(a).all())
