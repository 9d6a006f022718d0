"""Generate tests/golden/*.npz by running the REAL reference (TEST INFRASTRUCTURE).

Run in the build container only (`python oracle/make_golden.py`): it imports the unmodified
reference from /root/reference (read-only), which does not exist on the GPU box.  The vectors it
writes are committed; tests read only the .npz files.

Shims: `kornia` is absent -> a stub module (only used by colour transfer,
optex.py:126-128); the rotation is injected by replacing `optex.random_rotation` with a function
returning a fixed matrix (the reference draws it from numpy's global RNG, optex.py:149).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    kornia = types.ModuleType("kornia")
    color = types.ModuleType("kornia.color")
    hls = types.ModuleType("kornia.color.hls")
    hls.hls_to_rgb = hls.rgb_to_hls = lambda x: (_ for _ in ()).throw(RuntimeError("kornia stub"))
    kornia.color, color.hls = color, hls
    sys.modules.update({"kornia": kornia, "kornia.color": color, "kornia.color.hls": hls})
    from PIL import Image

    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    sys.path.insert(0, REF)
    import histmatch
    import optex
    import util

    return histmatch, optex, util


def features(seed, shape_t, shape_s, quant=False):
    """SURVEY.md §8(c) recipe."""
    g = torch.Generator().manual_seed(seed)
    t = torch.relu(torch.randn(*shape_t, generator=g))
    s = torch.relu(1.5 * torch.randn(*shape_s, generator=g) + 0.25)
    if quant:  # colour-like data with massive ties (H8)
        t = torch.round(t.clamp(0, 1) * 255) / 255
        s = torch.round(s.clamp(0, 1) * 255) / 255
    return t, s


def write_vgg_golden():
    """vgg.Encoder / vgg.Decoder of the reference with its own .pth weights (vgg.py:138-171) on a small image."""
    import vgg  # noqa: E402  (reference module, sys.path set by import_reference)

    x = torch.rand(2, 3, 21, 30, generator=torch.Generator().manual_seed(7))
    out = {"x": x.numpy()}
    with torch.inference_mode():
        for d in (1, 2, 3):
            f = vgg.Encoder(d)(x)
            out[f"enc{d}"] = f.numpy()
            out[f"dec{d}"] = vgg.Decoder(d)(f).numpy()
    np.savez_compressed(os.path.join(OUT, "vgg.npz"), **out)


UTIL_SIZE_CASES = [(512, 1.0, 736, 512, False), (512, 1.0, 736, 512, True), (1024, 1.0, 402, 402, True),
                   (2048, 1.0, 1920, 1088, False), (2048, 1.0, 3840, 2160, True), (256, 0.5, 300, 451, False),
                   (320, 2.0, 100, 77, True), (64, 1.0, 80, 64, False)]
UTIL_NAME_CASES = [
    dict(style=["style/graffiti.jpg"], content=None, mixing_alpha=0.5, content_strength=0.01, hist_mode="chol",
         no_pca=False, no_multires=False, style_scale=1.0, color_transfer=None, size=512, output_dir="output/"),
    dict(style=["style/zebra.jpg", "a/b/pattern-small.jpg"], content="content/rocket.jpg", mixing_alpha=0.25,
         content_strength=0.2, hist_mode="cdf", no_pca=True, no_multires=True, style_scale=0.5, color_transfer="opt",
         size=1024, output_dir="out"),
]


def write_util_golden(util):
    """util.get_size (util.py:33-42) and the file names util.save_image builds (util.py:45-65)."""
    import json
    from argparse import Namespace

    import torchvision

    out = {"get_size": [[list(c), list(util.get_size(*c))] for c in UTIL_SIZE_CASES], "names": []}
    real = torchvision.utils.save_image
    try:
        for batch in (1, 2):
            for case in UTIL_NAME_CASES:
                paths = []
                torchvision.utils.save_image = lambda t, path, *a, **k: paths.append(path)
                util.save_image(torch.zeros(batch, 3, 4, 4), Namespace(**case))
                out["names"].append({"batch": batch, "args": case, "paths": paths})
    finally:
        torchvision.utils.save_image = real
    with open(os.path.join(OUT, "util.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def write_texture_golden(optex):
    """optex.OptimalTexture.forward (optex.py:81-139) of the real reference, with seeded random weights in place of
    ./models/*.pth (too large to commit), the rotation stream of `texture_rotation`, torch's RNG seeded for the
    mixing mask, and oracle/image_oracle's HLS restatement standing in for the absent kornia (so the colour-transfer
    ORCHESTRATION is the reference's while the HLS arithmetic stays unpinned)."""
    import vgg  # reference module

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import image_oracle, vgg_oracle
    from oracle.texture_cases import TEXTURE_CASES, texture_inputs, texture_rotation

    def fake_load(path, *a, **k):
        base = os.path.basename(path)
        depth = int(base.split("conv")[1][0])
        sd = vgg_oracle.random_state_dict("encoder" if base.startswith("vgg_normalised") else "decoder", depth)
        keys = list(real_load(path, *a, **k).keys())       # the Sequential's own parameter names, in order
        assert len(keys) == len(sd)
        return dict(zip(keys, sd.values()))

    real_load = vgg.torch.load
    vgg.torch.load = fake_load
    optex.rgb_to_hls, optex.hls_to_rgb = image_oracle.rgb_to_hls, image_oracle.hls_to_rgb
    out = {}
    try:
        for name in TEXTURE_CASES:
            kwargs, styles, content, pastiche = texture_inputs(name)
            calls = [0]

            def rot(n, device="cpu", impl="scipy", _c=calls):
                r = texture_rotation(n, _c[0])
                _c[0] += 1
                return r

            optex.random_rotation = rot
            with torch.inference_mode():
                model = optex.OptimalTexture(**kwargs)       # nn.Conv2d initialisers consume torch's RNG
                torch.manual_seed(77)                        # -> the mixing-mask noise (optex.py:98) is torch.rand after this
                res = model.forward(pastiche.clone(), [s.clone() for s in styles],
                                    None if content is None else content.clone())
            out[f"{name}_out"] = res.contiguous().numpy()
            out[f"{name}_calls"] = np.asarray(calls[0])
    finally:
        vgg.torch.load = real_load
    np.savez_compressed(os.path.join(OUT, "texture.npz"), **out)


def main():
    histmatch, optex, util = import_reference()
    if "--texture-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        write_texture_golden(optex)
        write_util_golden(util)
        return
    from scipy.stats import special_ortho_group

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)

    # ---- interp / cdf_match known answers + random cases
    kat = {}
    f = lambda *v: torch.tensor(v, dtype=torch.float32)
    kat["interp0_x"], kat["interp0_xp"], kat["interp0_fp"] = f(.1, .5, .75, 1.), f(0, .5, 1.), f(10, 20, 40)
    kat["interp1_x"], kat["interp1_xp"], kat["interp1_fp"] = f(.2, .5, .6), f(0, .5, .5, 1.), f(1, 2, 3, 4)
    for k in (0, 1):
        kat[f"interp{k}_out"] = histmatch.interp(kat[f"interp{k}_x"], kat[f"interp{k}_xp"], kat[f"interp{k}_fp"])
    kat["cdf0_t"] = torch.arange(8, dtype=torch.float32)[None]
    kat["cdf0_s"] = f(10, 10, 12, 14, 14, 14, 18, 20)[None]
    kat["cdf0_out"] = histmatch.cdf_match(kat["cdf0_t"], kat["cdf0_s"])
    kat["cdf1_t"] = torch.full((1, 9), 3.25)  # constant channel, hi == lo
    kat["cdf1_s"] = torch.full((1, 5), 3.25)
    kat["cdf1_out"] = histmatch.cdf_match(kat["cdf1_t"], kat["cdf1_s"])
    g = torch.Generator().manual_seed(7)
    kat["cdf2_t"] = torch.randn(6, 1500, generator=g) * 2.0
    kat["cdf2_s"] = torch.relu(torch.randn(6, 1111, generator=g) * 1.3 + 0.2)
    kat["cdf2_out"] = histmatch.cdf_match(kat["cdf2_t"], kat["cdf2_s"])
    np.savez_compressed(os.path.join(OUT, "kat_interp_cdf.npz"), **{k: v.numpy() for k, v in kat.items()})

    # ---- hist_match without rotation (SURVEY.md §8c sanity pin, seed 123)
    hm = {}
    t, s = features(123, (1, 8, 8, 16), (1, 6, 10, 16))
    hm["t"], hm["s"] = t, s
    for mode in ("chol", "pca", "sym", "cdf"):
        hm[f"out_{mode}"] = histmatch.hist_match(t, s, mode).contiguous()
    np.savez_compressed(os.path.join(OUT, "hist_match_seed123.npz"), **{k: v.numpy() for k, v in hm.items()})

    # ---- optimal_transport with an injected rotation
    cases = {
        # name: (seed, target shape, source shape, quantised)
        "sq16": (1, (1, 8, 8, 16), (1, 8, 8, 16), False),
        "ragged23": (2, (1, 12, 10, 23), (1, 9, 14, 23), False),   # C after PCA is arbitrary (H6), N_s != N_p
        "batch2": (3, (2, 6, 6, 8), (2, 5, 7, 8), False),          # batch coupling (H7), b_s == b
        "batch2_s1": (4, (2, 6, 6, 8), (1, 8, 8, 8), False),       # b_s == 1 broadcast (histmatch.py:44)
        "rgb_ties": (5, (1, 24, 24, 3), (1, 20, 28, 3), True),     # colour transfer shape (optex.py:131-133)
        "wide64": (6, (1, 16, 16, 64), (1, 16, 16, 64), False),
    }
    ot = {}
    for name, (seed, st, ss, quant) in cases.items():
        t, s = features(seed, st, ss, quant)
        c = st[-1]
        rot = torch.tensor(special_ortho_group.rvs(c, random_state=seed))  # float64, as optex.py:149
        optex.random_rotation = lambda n, device="cpu", impl="scipy", _r=rot: _r
        ot[f"{name}_t"], ot[f"{name}_s"], ot[f"{name}_rot"] = t, s, rot
        modes = ("cdf",) if name == "rgb_ties" else ("chol", "pca", "sym", "cdf")
        for mode in modes:
            ot[f"{name}_out_{mode}"] = optex.optimal_transport(t, s, mode).contiguous()
    np.savez_compressed(os.path.join(OUT, "ot_step.npz"), **{k: v.numpy() for k, v in ot.items()})

    # ---- inner loop with content blend (optex.py:112-117), 3 iterations, l = 1 (conv4_1)
    lp = {}
    t, s = features(11, (1, 8, 8, 16), (1, 8, 8, 16))
    content = torch.relu(torch.randn(1, 8, 8, 16, generator=torch.Generator().manual_seed(12)))
    rots = [torch.tensor(special_ortho_group.rvs(16, random_state=100 + i)) for i in range(3)]
    lp["t"], lp["s"], lp["content"], lp["rots"] = t, s, content, torch.stack(rots)
    for mode in ("chol", "cdf"):
        p = t.clone()
        for r in rots:
            optex.random_rotation = lambda n, device="cpu", impl="scipy", _r=r: _r
            p = optex.optimal_transport(p, s, mode)
            strength = 0.2 / 2 ** (4 - 1)
            p += strength * (content - p)
        lp[f"out_{mode}"] = p.contiguous()
    np.savez_compressed(os.path.join(OUT, "inner_loop.npz"), **{k: v.numpy() for k, v in lp.items()})

    # ---- scipy rotation stream + fit_pca + schedule (host logic)
    misc = {}
    for n in (3, 16, 23):
        misc[f"so_{n}_seed5"] = special_ortho_group.rvs(n, random_state=5)
    x = torch.relu(torch.randn(1, 12, 12, 32, generator=torch.Generator().manual_seed(21))) * torch.linspace(0.1, 3, 32)
    feats, eig = optex.fit_pca(x)
    misc["pca_x"], misc["pca_feats"], misc["pca_eig"] = x.numpy(), feats.numpy(), eig.numpy()
    for (size, iters, passes) in ((512, 500, 5), (256, 500, 4), (1024, 500, 5), (2048, 300, 3)):
        its, sizes = util.get_iters_and_sizes(size, iters, passes, True)
        misc[f"sched_{size}_{iters}_{passes}_iters"] = np.asarray(its)
        misc[f"sched_{size}_{iters}_{passes}_sizes"] = np.asarray(sizes)
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **misc)
    write_vgg_golden()
    torch.set_num_threads(os.cpu_count() or 1)
    write_texture_golden(optex)
    write_util_golden(util)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
