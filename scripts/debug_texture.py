"""Stage-by-stage comparison of the synthesis loop (B200 vs CPU oracle), each GPU stage fed the ORACLE's inputs so
errors do not compound; every GPU stage is run twice to expose non-determinism.  Test infrastructure."""
import sys

import torch

sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from optimaltextures_b200 import texture, util, vgg
from oracle import image_oracle, ot_oracle, texture_cases, texture_oracle, vgg_oracle

name = sys.argv[1] if len(sys.argv) > 1 else "synth_pca"
if name == "nomultires":
    kwargs = dict(size=64, iters=10, passes=2, hist_mode="chol", no_multires=True)
    g0 = torch.Generator().manual_seed(5)
    styles, content, pastiche = [torch.rand(1, 3, 64, 96, generator=g0)], None, torch.rand(1, 3, 64, 64, generator=g0)
else:
    kwargs, styles, content, pastiche = texture_cases.texture_inputs(name)
sd = texture_cases.state_dicts()
mode = kwargs["hist_mode"]
use_pca = not kwargs.get("no_pca", False)
its, sizes = texture_oracle.get_iters_and_sizes(kwargs["size"], kwargs["iters"], kwargs["passes"],
                                                 not kwargs.get("no_multires", False))
cs = kwargs.get("content_strength", 0.1)
encs = {d: vgg.Encoder(d, state_dict=sd[("encoder", d)]) for d in range(1, 6)}
decs = {d: vgg.Decoder(d, state_dict=sd[("decoder", d)]) for d in range(1, 6)}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def twice(fn):
    x, y = fn(), fn()
    torch.cuda.synchronize()
    same = all(torch.equal(a, b) for a, b in zip(x, y)) if isinstance(x, (list, tuple)) else torch.equal(x, y)
    return x, same


calls = 0
for p, size in enumerate(sizes):
    print(f"== pass {p} size {size}", flush=True)
    if pastiche.shape[-2] != size and pastiche.shape[-1] != size:
        ssz = texture_oracle.get_size(size, 1, styles[0].shape[2], styles[0].shape[3])
        st_ref = image_oracle.resize(styles[0], ssz)
        st_gpu, same = twice(lambda: util.resize(styles[0].cuda(), ssz))
        print(f"resize style {tuple(styles[0].shape)} -> {ssz}: err {rel(st_gpu, st_ref):.2e} deterministic {same}")
        pa_in = pastiche
        pa_ref = image_oracle.resize(pastiche, (size, size))
        pa_gpu, same = twice(lambda: util.resize(pastiche.cuda(), (size, size)))
        print(f"resize pastiche -> {size}: err {rel(pa_gpu, pa_ref):.2e} deterministic {same}")
        pastiche, style_t = pa_ref, st_ref
        cont_t = None
        if content is not None:
            csz = texture_oracle.get_size(size, 1.0, content.shape[2], content.shape[3], oversize=True)
            cont_t = image_oracle.resize(content, csz)
            pastiche = image_oracle.resize(pa_in, csz)
    else:
        style_t, cont_t = styles[0], content
    all_gpu, same = twice(lambda: encs[5].forward_all(style_t.cuda()))
    print(f"forward_all(style) deterministic {same}")
    for l, d in enumerate(range(5, 0, -1)):
        sf = vgg_oracle.encoder_forward(style_t, sd[("encoder", d)], d)
        print(f"-- layer conv{d}_1 style feat {tuple(sf.shape)}: forward_all err {rel(all_gpu[d - 1], sf):.2e}", flush=True)
        eig = None
        if use_pca:
            (fg, eg), same = twice(lambda: ob.fit_pca(sf.cuda()))
            _, e_ref = ot_oracle.fit_pca(sf)
            proj = rel(eg @ eg.T, e_ref @ e_ref.T) if eg.shape == e_ref.shape else float("nan")
            print(f"   fit_pca k gpu {eg.shape[1]} ref {e_ref.shape[1]} projector err {proj:.2e} feats-vs-(x@eig) "
                  f"{rel(fg, sf @ eg.cpu()):.2e} deterministic {same} finite {bool(torch.isfinite(eg).all())}")
            eig = eg.cpu()
            sf = sf @ eig
        f_ref = vgg_oracle.encoder_forward(pastiche, sd[("encoder", d)], d)
        f_gpu, same = twice(lambda: encs[d](pastiche.cuda()))
        print(f"   encode pastiche err {rel(f_gpu, f_ref):.2e} deterministic {same}")
        if use_pca:
            pg, same = twice(lambda: ob.pca_project(f_ref.cuda(), eig.cuda()))
            f_ref = f_ref @ eig
            print(f"   project err {rel(pg, f_ref):.2e} deterministic {same}")
        cf = None
        if cont_t is not None:
            cf = vgg_oracle.encoder_forward(cont_t, sd[("encoder", d)], d)
            if use_pca:
                cf = cf @ eig
            cf_ref = image_oracle.recentre(cf, sf)
            cf_gpu, same = twice(lambda: texture.recentre(cf.cuda(), sf.cuda()))
            print(f"   recentre err {rel(cf_gpu, cf_ref):.2e} deterministic {same}")
            cf = cf_ref if l <= 2 else None
        strength = cs / 2 ** (4 - l)
        n = its[p][l - 1]
        rots = [texture_cases.texture_rotation(f_ref.shape[-1], calls + i) for i in range(n)]
        calls += n
        if n:
            o_ref = ot_oracle.ot_loop(f_ref.clone(), sf, rots, mode, content=cf, content_strength=cs, l=l)
            rg = torch.stack([r.float() for r in rots]).cuda()
            o_gpu, same = twice(lambda: ob.ot_loop(f_ref.cuda(), sf.cuda(), mode, n, rotations=rg,
                                                   content=None if cf is None else cf.cuda(),
                                                   content_strength=strength if cf is not None else 0.0))
            o_dev, same2 = twice(lambda: ob.ot_loop(f_ref.cuda(), sf.cuda(), mode, n, seed=3, first_counter=0))
            step, _ = twice(lambda: ob.optimal_transport(f_ref.cuda(), sf.cuda(), mode, rotation=rg[0]))
            s_ref = ot_oracle.ot_step(f_ref, sf, rots[0], mode)
            print(f"   finite gpu {bool(torch.isfinite(o_gpu).all())} ref {bool(torch.isfinite(o_ref).all())} |ref|max {float(o_ref.abs().max()):.3g}")
            print(f"   ot_loop x{n} c={f_ref.shape[-1]} n_p={f_ref.shape[1] * f_ref.shape[2]} n_s={sf.shape[1] * sf.shape[2]}: "
                  f"err {rel(o_gpu, o_ref):.2e} deterministic {same} (device rotations: {same2}); single step err "
                  f"{rel(step, s_ref):.2e}", flush=True)
            f_ref = o_ref
        if use_pca:
            ug, same = twice(lambda: ob.pca_project(f_ref.cuda(), eig.cuda(), transpose=True))
            f_ref = f_ref @ eig.T
            print(f"   unproject err {rel(ug, f_ref):.2e} deterministic {same}")
        d_ref = vgg_oracle.decoder_forward(f_ref, sd[("decoder", d)], d)
        d_gpu, same = twice(lambda: decs[d](f_ref.cuda()))
        print(f"   decode err {rel(d_gpu, d_ref):.2e} deterministic {same}", flush=True)
        pastiche = d_ref
