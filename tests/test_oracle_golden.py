"""Pin the oracle (oracle/*.py) against vectors produced by the REAL reference
(tests/golden/*.npz, written by oracle/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import cdf_explicit, ot_oracle, rotation

T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
OT_CASES = ["sq16", "ragged23", "batch2", "batch2_s1", "wide64"]


def test_interp_known_answers(golden):
    g = golden("kat_interp_cdf")
    np.testing.assert_array_equal(g["interp0_out"], np.float32([4, 20, 40, 40]))
    np.testing.assert_array_equal(g["interp1_out"], np.float32([2, 2, 4]))
    for k in (0, 1):
        args = [g[f"interp{k}_{n}"] for n in ("x", "xp", "fp")]
        np.testing.assert_array_equal(ot_oracle.interp_backward(*map(T, args)).numpy(), g[f"interp{k}_out"])
        np.testing.assert_array_equal(cdf_explicit.interp_backward(*args), g[f"interp{k}_out"])


@pytest.mark.parametrize("k", [0, 1, 2])
def test_cdf_match_bit_exact(golden, k):
    g = golden("kat_interp_cdf")
    t, s, ref = g[f"cdf{k}_t"], g[f"cdf{k}_s"], g[f"cdf{k}_out"]
    np.testing.assert_array_equal(ot_oracle.cdf_match_channels(T(t), T(s)).numpy(), ref)
    np.testing.assert_array_equal(cdf_explicit.cdf_match(t, s), ref)
    if k == 0:
        np.testing.assert_array_equal(
            ref[0], np.float32([10.078125, 10.078125, 12.03125, 14.0625, 14.0625, 14.0625, 18.046875, 20.0]))


@pytest.mark.parametrize("mode", ["chol", "pca", "sym", "cdf"])
def test_hist_match_seed123(golden, mode):
    g = golden("hist_match_seed123")
    out = ot_oracle.hist_match_nhwc(T(g["t"]), T(g["s"]), mode).contiguous().numpy()
    np.testing.assert_array_equal(out, g[f"out_{mode}"])
    pins = {"chol": 762.895552, "pca": 762.895551, "sym": 762.895550, "cdf": 961.213475}
    assert abs(float(out.sum(dtype=np.float64)) - pins[mode]) < 2e-6


@pytest.mark.parametrize("name", OT_CASES + ["rgb_ties"])
def test_ot_step_bit_exact(golden, name):
    g = golden("ot_step")
    modes = ("cdf",) if name == "rgb_ties" else ("chol", "pca", "sym", "cdf")
    for mode in modes:
        out = ot_oracle.ot_step(T(g[f"{name}_t"]), T(g[f"{name}_s"]), T(g[f"{name}_rot"]), mode)
        np.testing.assert_array_equal(out.contiguous().numpy(), g[f"{name}_out_{mode}"])


@pytest.mark.parametrize("name", OT_CASES + ["rgb_ties"])
def test_explicit_cdf_inside_ot_step(golden, name):
    """numpy-explicit cdf on the oracle's rotated features == reference cdf step, bit for bit."""
    g = golden("ot_step")
    t, s, rot = T(g[f"{name}_t"]), T(g[f"{name}_s"]), T(g[f"{name}_rot"]).float()
    c = t.shape[-1]
    rp, rs = (t @ rot).reshape(-1, c).T.contiguous(), (s @ rot).reshape(-1, c).T.contiguous()
    m = T(cdf_explicit.cdf_match(rp.numpy(), rs.numpy())).T.reshape(t.shape)
    np.testing.assert_array_equal((m @ rot.T).numpy(), g[f"{name}_out_cdf"])


@pytest.mark.parametrize("mode", ["chol", "cdf"])
def test_inner_loop_with_content(golden, mode):
    g = golden("inner_loop")
    out = ot_oracle.ot_loop(T(g["t"]).clone(), T(g["s"]), list(T(g["rots"])), mode,
                            content=T(g["content"]), content_strength=0.2, l=1)
    np.testing.assert_array_equal(out.contiguous().numpy(), g[f"out_{mode}"])


def test_fit_pca(golden):
    g = golden("misc")
    feats, eig = ot_oracle.fit_pca(T(g["pca_x"]))
    np.testing.assert_array_equal(eig.numpy(), g["pca_eig"])
    np.testing.assert_array_equal(feats.numpy(), g["pca_feats"])


@pytest.mark.parametrize("n", [3, 16, 23])
def test_scipy_rotation_restated(golden, n):
    g = golden("misc")
    np.testing.assert_array_equal(rotation.haar_rotation_qr(n, 5), g[f"so_{n}_seed5"])
    from scipy.stats import special_ortho_group      # installed 1.18.1; the reference's call (optex.py:149)
    np.testing.assert_array_equal(rotation.haar_rotation_qr(n, 9), special_ortho_group.rvs(n, random_state=9))


def test_householder_rotation_is_special_orthogonal():
    rng = np.random.RandomState(0)
    for n in (2, 5, 32):
        h = rotation.haar_rotation_householder(rng.normal(size=(n - 1, n)))
        np.testing.assert_allclose(h @ h.T, np.eye(n), atol=1e-12)
        assert abs(np.linalg.det(h) - 1) < 1e-10


# ---------------------------------------------------------------------------------------------- synthesis loop
@pytest.mark.parametrize("name", ["synth_pca", "mix_content_chol_opt", "nopca_cdf_lum"])
def test_texture_forward_matches_reference(golden, name):
    """oracle/texture_oracle.OptimalTexture == the real reference's OptimalTexture.forward (optex.py:81-139) on the
    same seeded weights / inputs / rotation stream / mask noise, bit for bit (same torch ops in the same order)."""
    from oracle import texture_cases, texture_oracle

    g = golden("texture")
    kwargs, styles, content, pastiche = texture_cases.texture_inputs(name)
    torch.manual_seed(77)
    model = texture_oracle.OptimalTexture(texture_cases.state_dicts(), rotation_fn=texture_cases.texture_rotation,
                                          **kwargs)
    with torch.inference_mode():
        out = model.forward(pastiche, styles, content)
    assert model.ot_calls == int(g[f"{name}_calls"])
    np.testing.assert_array_equal(out.contiguous().numpy(), g[f"{name}_out"])


def test_schedule_and_sizes(golden):
    from oracle import texture_oracle

    g = golden("misc")
    for (size, iters, passes) in ((512, 500, 5), (256, 500, 4), (1024, 500, 5), (2048, 300, 3)):
        its, sizes = texture_oracle.get_iters_and_sizes(size, iters, passes, True)
        np.testing.assert_array_equal(np.asarray(its), g[f"sched_{size}_{iters}_{passes}_iters"])
        np.testing.assert_array_equal(np.asarray(sizes), g[f"sched_{size}_{iters}_{passes}_sizes"])
