"""CPU tests for the VGG row: the oracle against the reference's golden outputs, and the host logic
of optimaltextures_b200.vgg (1x1 fold, weight packing order, layer tables) against the oracle through a pure-torch
emulation of what the CUDA gather + GEMM compute (no compute call into the library: there is no GPU here)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vgg_oracle

T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
REF_MODELS = os.environ.get("OPTEX_MODELS_DIR", "/root/reference/models")


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="needs the reference's .pth weights (build container only)")
@pytest.mark.parametrize("depth", [1, 2, 3])
def test_oracle_matches_reference_golden(golden, depth):
    g = golden("vgg")
    enc = torch.load(os.path.join(REF_MODELS, f"vgg_normalised_conv{depth}_1.pth"), map_location="cpu")
    dec = torch.load(os.path.join(REF_MODELS, f"feature_invertor_conv{depth}_1.pth"), map_location="cpu")
    # same torch ops as the reference's modules, but torch's CPU conv picks its summation order by thread count
    # (the fixture was written single-threaded; run in one process with equal settings the two are bit-identical):
    # tolerance = a few fp32 ulps of the tensor's scale
    f = vgg_oracle.encoder_forward(T(g["x"]), enc, depth)
    assert float(np.abs(f.numpy() - g[f"enc{depth}"]).max()) <= 4e-6 * float(np.abs(g[f"enc{depth}"]).max())
    img = vgg_oracle.decoder_forward(T(g[f"enc{depth}"]), dec, depth).numpy()
    assert float(np.abs(img - g[f"dec{depth}"]).max()) <= 4e-6 * max(1.0, float(np.abs(g[f"dec{depth}"]).max()))


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="needs the reference's .pth weights (build container only)")
def test_encoder_files_are_prefix_identical():
    """Encoder(5)'s first layers equal Encoder(d)'s - the basis of Encoder.forward_all."""
    big = list(torch.load(os.path.join(REF_MODELS, "vgg_normalised_conv5_1.pth"), map_location="cpu").values())
    for d in (1, 2, 3, 4):
        small = list(torch.load(os.path.join(REF_MODELS, f"vgg_normalised_conv{d}_1.pth"), map_location="cpu").values())
        assert all(torch.equal(a, b) for a, b in zip(small, big))


def emulate_layer(x_nhwc, layer):
    """what optex_conv3x3 computes, restated with torch ops on the PACKED operands (tap-major im2col @ W^T)"""
    x = x_nhwc.permute(0, 3, 1, 2)
    if layer.pre == 1:
        x = F.max_pool2d(x, 2, 2, 0, ceil_mode=True)
    elif layer.pre == 2:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    b, c, h, w = x.shape
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    taps = [xp[:, :, ky:ky + h, kx:kx + w].permute(0, 2, 3, 1) for ky in range(3) for kx in range(3)]
    col = torch.cat(taps, dim=-1).reshape(b * h * w, 9 * c)
    kp = layer.w.shape[1]
    col = F.pad(col, (0, kp - 9 * c))
    out = col.double() @ layer.w.double().T + layer.b.double()
    if layer.relu:
        out = torch.relu(out)
    return out.float().reshape(b, h, w, -1)


@pytest.mark.parametrize("depth", [1, 3, 5])
def test_host_packing_and_fold_reproduce_the_oracle(depth):
    from optimaltextures_b200 import vgg

    torch.manual_seed(depth)
    x = torch.rand(1, 3, 18, 23)
    sd_e = vgg_oracle.random_state_dict("encoder", depth, seed=depth)
    sd_d = vgg_oracle.random_state_dict("decoder", depth, seed=10 + depth)
    enc = vgg.Encoder(depth, state_dict=sd_e, device="cpu")        # packing only: nothing is computed by the library
    dec = vgg.Decoder(depth, state_dict=sd_d, device="cpu")
    cur = x.permute(0, 2, 3, 1)
    for layer in enc.layers:
        cur = emulate_layer(cur, layer)
    ref = vgg_oracle.encoder_forward(x, sd_e, depth)
    assert cur.shape == ref.shape
    assert float((cur - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    for layer in dec.layers:
        cur = emulate_layer(cur, layer)
    img = cur[..., :3].permute(0, 3, 1, 2)
    ref_img = vgg_oracle.decoder_forward(ref, sd_d, depth)
    assert img.shape == ref_img.shape
    assert float((img - ref_img).abs().max()) <= 5e-5 * max(1.0, float(ref_img.abs().max()))
    assert dec.layers[-1].cout_pad == 32 and float(dec.layers[-1].w[3:].abs().max()) == 0.0


def test_layer_tables_agree_with_the_oracle():
    from optimaltextures_b200 import vgg

    assert vgg._ENCODER == vgg_oracle.ENCODER_CONVS and vgg._ENCODER_END == vgg_oracle.ENCODER_DEPTH_END
    assert vgg._DECODER == vgg_oracle.DECODER_BLOCKS


def test_no_cpu_fallback():
    from optimaltextures_b200 import vgg

    enc = vgg.Encoder(1, state_dict=vgg_oracle.random_state_dict("encoder", 1), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.rand(1, 3, 8, 8))


def test_random_state_dicts_match_the_golden_runs_weights():
    """optimaltextures_b200.vgg.random_state_dicts (what bench.py uses: the product path must not import oracle/) draws
    the same seeded weights as the oracle's generator, in the reference's state_dict order (vgg.py:14-136)."""
    from optimaltextures_b200 import vgg
    from oracle import texture_cases

    ours, ref = vgg.random_state_dicts(0), texture_cases.state_dicts(0)
    assert ours.keys() == ref.keys()
    for key in ours:
        assert list(ours[key].keys()) == list(ref[key].keys()), key
        for name in ours[key]:
            assert torch.equal(ours[key][name], ref[key][name]), (key, name)
