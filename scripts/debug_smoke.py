"""Finds the first stage of a tiny synthesis (the one __graft_entry__.smoke() runs) that produces non-finite values.
Test infrastructure: wraps the stage functions of optimaltextures_b200.texture with finiteness checks."""
import sys

import torch

sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from optimaltextures_b200 import optex as _optex, texture, vgg

sd = vgg.random_state_dicts(0)
g = torch.Generator().manual_seed(0)
_ = torch.relu(torch.randn(1, 32, 32, 64, generator=g)); _ = torch.randn(1, 24, 40, 64, generator=g)
img = torch.rand(1, 3, 32, 48, generator=g)
past = torch.rand(1, 3, 32, 32, generator=g)


def stats(name, t):
    t = t.float()
    fin = torch.isfinite(t)
    ok = bool(fin.all())
    print(f"    {name:28s} {str(tuple(t.shape)):22s} finite={ok} "
          f"absmax={float(t[fin].abs().max()) if fin.any() else float('nan'):.4g} bad={int((~fin).sum())}", flush=True)
    return ok


def wrap(mod, fname, label):
    orig = getattr(mod, fname)

    def f(*a, **k):
        out = orig(*a, **k)
        torch.cuda.synchronize()
        if isinstance(out, torch.Tensor):
            stats(label, out)
        elif isinstance(out, (list, tuple)):
            for i, o in enumerate(out):
                if isinstance(o, torch.Tensor):
                    stats(f"{label}[{i}]", o)
                elif isinstance(o, (list, tuple)):
                    for j, oo in enumerate(o):
                        if isinstance(oo, torch.Tensor):
                            stats(f"{label}[{i}][{j}]", oo)
        return out

    setattr(mod, fname, f)
    return orig


origs = [(m, n, wrap(m, n, n)) for m, n in ((_optex, "fit_pca_many"), (_optex, "pca_project"), (_optex, "ot_loop"),
                                           (texture._util, "resize"))]
dec_fwd = vgg.Decoder.forward
enc_fwd = vgg.Encoder.forward


def dfw(self, x):
    out = dec_fwd(self, x); torch.cuda.synchronize(); stats(f"Decoder({self.depth})", out); return out


def efw(self, x):
    out = enc_fwd(self, x); torch.cuda.synchronize(); stats(f"Encoder({self.depth})", out); return out


vgg.Decoder.forward = dfw; vgg.Decoder.__call__ = dfw
vgg.Encoder.forward = efw; vgg.Encoder.__call__ = efw

args = sys.argv[1:]
PRELUDE = "--prelude" in args
INSTRUMENT = "--instrument" in args
variants = [a for a in args if not a.startswith("--")] or ["default"]
if not INSTRUMENT:
    for m, n, o in origs:
        setattr(m, n, o)
    vgg.Decoder.forward = dec_fwd; vgg.Decoder.__call__ = dec_fwd
    vgg.Encoder.forward = enc_fwd; vgg.Encoder.__call__ = enc_fwd
if "--smoke" in args:
    import __graft_entry__ as ge
    ge.smoke()
    sys.exit(0)
if PRELUDE:                                     # what smoke() runs before the synthesis
    g2 = torch.Generator().manual_seed(0)
    p = torch.relu(torch.randn(1, 32, 32, 64, generator=g2))
    s_ = torch.relu(1.3 * torch.randn(1, 24, 40, 64, generator=g2) + 0.2)
    r = ob.random_rotation(64, "cuda", seed=0, counter=0)
    rp, rs = ob.rotate_forward(p.cuda(), r), ob.rotate_forward(s_.cuda(), r)
    ob.cdf_match(rp, rs)
    pre = [a for a in args if a.startswith("--modes=")]
    modes = pre[0][8:].split(",") if pre else ["cdf", "sort", "chol", "pca", "sym"]
    for mode in modes:
        if mode:
            ob.optimal_transport(p.cuda(), s_.cuda(), mode, rotation=r)
    if "--novgg" not in args:
        feat = vgg.Encoder(3, state_dict=sd[("encoder", 3)])(img.cuda())
        vgg.Decoder(3, state_dict=sd[("decoder", 3)])(feat)
    print("prelude done", flush=True)
for v in variants:
    print(f"=== variant {v}", flush=True)
    kw = dict(size=32, iters=10, passes=1, hist_mode="pca", state_dicts=sd)
    if v == "chol":
        kw["hist_mode"] = "chol"
    if v == "size64":
        kw.update(size=64, iters=24, passes=2)
    model = texture.OptimalTexture(**kw)
    if v == "nopad":
        model.pad_channels = 1
    if v == "seqpca":
        model.pca = lambda t: _optex.fit_pca(t)
    ob.manual_seed(0)
    out = model.forward(past.cuda(), [img.cuda()])
    torch.cuda.synchronize()
    print(f"  result finite={bool(torch.isfinite(out).all())} k={model.last_pca_k}", flush=True)
