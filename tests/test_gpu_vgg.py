"""GPU parity: the VGG-19 encoder / feature-inverter layers (vgg.py:14-171) through the C-ABI (`optex_conv3x3`,
`optex_nhwc_to_nchw`) vs the torch fp32 CPU oracle (oracle/vgg_oracle.py, pinned to the reference's own modules
and weights in tests/test_vgg_host.py).

Stated tolerance (floating point), |out - ref| <= tol * max(1, |ref|_max):
  * gemm mode "fp32" (FFMA tiles, round-to-nearest fp32 accumulation in a different order from torch's CPU conv):
    2e-5 per layer, 2e-4 through a whole 13-layer encoder / decoder stack;
  * gemm mode "auto" (3xTF32 on the tensor cores): the products are fp32-grade, but tcgen05 accumulates in fp32
    with truncation, so the error grows linearly with the reduction length K = 9 c_in: measured 4e-6 at K = 512
    (the rotations) and 4e-5 at K = 4608 (conv4_x / conv5_1).  1e-4 per layer, 1e-3 through a stack - still ~10x
    tighter than the reference's own CUDA path (cuDNN convs with TF32 inputs, optex.py:248-249)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vgg_oracle

pytestmark = pytest.mark.gpu
LAYER_TOL = {"fp32": 2e-5, "auto": 1e-4}
STACK_TOL = {"fp32": 2e-4, "auto": 1e-3}


@pytest.fixture(scope="module")
def ob():
    import optimaltextures_b200 as ob

    return ob


@pytest.fixture(params=["auto", "fp32"])
def gemm_mode(request, ob):
    ob.set_gemm_mode(request.param)
    yield request.param
    ob.set_gemm_mode("auto")


def close(out, ref, tol):
    out, ref = out.double().cpu(), ref.double()
    assert out.shape == ref.shape, f"{tuple(out.shape)} vs {tuple(ref.shape)}"
    err = float((out - ref).abs().max())
    lim = tol * max(1.0, float(ref.abs().max()))
    assert err <= lim, f"max |err| {err:.3e} > {lim:.3e}"


def ref_layer(x_nchw, w, b, pre, relu):
    return vgg_oracle._layer(x_nchw, pre, w, b, relu).permute(0, 2, 3, 1).contiguous()


LAYERS = [  # b, h_src, w_src, c_in, c_out, pre, relu
    (1, 13, 17, 3, 64, 0, True), (2, 9, 12, 64, 128, 1, True), (1, 7, 5, 128, 64, 2, True), (1, 32, 32, 64, 3, 0, False),
    (1, 16, 16, 512, 512, 0, True), (1, 33, 17, 256, 512, 1, True), (2, 4, 6, 512, 256, 2, True), (1, 2, 2, 64, 64, 0, True),
    (1, 3, 5, 20, 24, 1, True), (1, 130, 140, 64, 64, 0, True),
]


@pytest.mark.parametrize("b,hs,ws,cin,cout,pre,relu", LAYERS)
def test_conv3x3_layer(ob, gemm_mode, b, hs, ws, cin, cout, pre, relu):
    from optimaltextures_b200 import vgg

    g = torch.Generator().manual_seed(hs * 100 + ws + cin)
    x = torch.randn(b, cin, hs, ws, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    ref = ref_layer(x, w, bias, pre, relu)
    out = vgg.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w, bias, pre, relu)
    close(out, ref, LAYER_TOL[gemm_mode])
    if cin == 3:                                     # the first layer can read the NCHW image directly
        out2 = vgg.conv3x3(x.cuda(), w, bias, pre, relu, src_nchw=True)
        close(out2, ref, LAYER_TOL[gemm_mode])


def test_conv3x3_chunked_rows_equal_one_pass(ob, monkeypatch):
    """rows are processed in L2-sized chunks: a layer larger than one chunk (64 MB of im2col here: 2 chunks)"""
    from optimaltextures_b200 import vgg

    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 64, 200, 150, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    bias = torch.randn(64, generator=g) * 0.1
    out = vgg.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w, bias, 0, True)
    close(out, ref_layer(x, w, bias, 0, True), LAYER_TOL["auto"])


@pytest.mark.parametrize("depth,b,h,w", [(1, 1, 24, 31), (2, 2, 21, 30), (3, 1, 40, 56), (5, 1, 67, 50), (5, 2, 32, 48)])
def test_encoder_decoder_stacks(ob, gemm_mode, depth, b, h, w):
    from optimaltextures_b200 import vgg

    x = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(depth * 7 + h))
    sd_e = vgg_oracle.random_state_dict("encoder", depth, seed=depth)
    sd_d = vgg_oracle.random_state_dict("decoder", depth, seed=20 + depth)
    enc, dec = vgg.Encoder(depth, state_dict=sd_e), vgg.Decoder(depth, state_dict=sd_d)
    ref_all = vgg_oracle.encoder_forward(x, sd_e, depth, all_depths=True)
    outs = enc.forward_all(x.cuda())
    assert len(outs) == depth
    for o, r in zip(outs, ref_all):
        assert o.is_contiguous()
        close(o, r, STACK_TOL[gemm_mode])
    feat = enc(x.cuda())
    assert torch.equal(feat, outs[-1])
    ref_img = vgg_oracle.decoder_forward(ref_all[-1], sd_d, depth)
    img = dec(ref_all[-1].cuda())                    # decoder on the ORACLE's features: isolates the decoder
    assert img.shape == ref_img.shape and img.is_contiguous()
    close(img, ref_img, STACK_TOL[gemm_mode])
    close(dec(feat), ref_img, 5 * STACK_TOL[gemm_mode])         # and end to end


def test_golden_features_decode(ob, golden):
    """the reference's own encoder outputs (tests/golden/vgg.npz) have the shapes this module produces"""
    from optimaltextures_b200 import vgg

    g = golden("vgg")
    x = torch.from_numpy(g["x"])
    for d in (1, 2, 3):
        sd = vgg_oracle.random_state_dict("encoder", d, seed=d)
        assert tuple(vgg.Encoder(d, state_dict=sd)(x.cuda()).shape) == g[f"enc{d}"].shape


def test_errors(ob):
    from optimaltextures_b200 import vgg

    with pytest.raises(ValueError):
        vgg.conv3x3(torch.zeros(1, 1, 1, 8).cuda(), torch.zeros(8, 8, 3, 3), None)      # 1x1: no reflection pad
    enc = vgg.Encoder(1, state_dict=vgg_oracle.random_state_dict("encoder", 1))
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 4, 8, 8).cuda())


@pytest.mark.parametrize("size,depth", [(256, 3), (192, 4)])
def test_encoder_layers_at_image_size(ob, size, depth):
    """Every layer of the encoder at a real pass size (256^2 = pass 0 of every multires run): tens of thousands of
    GEMM rows, i.e. several im2col chunks per layer and several tiles per CTA - each layer fed the oracle's input,
    run three times: parity per layer and bit-identical repeats."""
    from optimaltextures_b200 import vgg

    sd = vgg_oracle.random_state_dict("encoder", depth)
    enc = vgg.Encoder(depth, state_dict=sd)
    x = torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(size))
    wb = vgg_oracle.pairs(sd)
    cur = F.conv2d(x, wb[0][0], wb[0][1])
    for i, ((pre, cin, cout, relu), (w, b)) in enumerate(zip(vgg_oracle.encoder_convs(depth), wb[1:])):
        ref = vgg_oracle._layer(cur, pre, w, b, relu)
        if i == 0:
            outs = [enc.layers[0].run(x.cuda(), src_nchw=True) for _ in range(3)]
        else:
            inp = cur.permute(0, 2, 3, 1).contiguous().cuda()
            outs = [enc.layers[i].run(inp) for _ in range(3)]
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2]), f"layer {i} is not deterministic"
        close(outs[0], ref.permute(0, 2, 3, 1), LAYER_TOL["auto"])
        cur = ref


def test_encoder_decoder_stack_256(ob):
    """Encoder(5) -> Decoder(5) at 256^2 vs the oracle (the chain every pass of the synthesis loop runs)."""
    from optimaltextures_b200 import vgg

    esd, dsd = vgg_oracle.random_state_dict("encoder", 5), vgg_oracle.random_state_dict("decoder", 5)
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    f_ref = vgg_oracle.encoder_forward(x, esd, 5)
    f = vgg.Encoder(5, state_dict=esd)(x.cuda())
    close(f, f_ref, STACK_TOL["auto"])
    img_ref = vgg_oracle.decoder_forward(f_ref, dsd, 5)
    img = vgg.Decoder(5, state_dict=dsd)(f_ref.cuda())
    close(img, img_ref, STACK_TOL["auto"])
