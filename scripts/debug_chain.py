"""Run the product synthesis loop with every stage wrapped to print finiteness / magnitude.  Test infrastructure."""
import sys

import torch

sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from optimaltextures_b200 import optex as _optex, texture
from oracle import texture_cases


def wrap(mod, name):
    fn = getattr(mod, name)

    def inner(*a, **k):
        out = fn(*a, **k)
        ts = out if isinstance(out, (tuple, list)) else (out,)
        torch.cuda.synchronize()
        info = " ".join(f"{tuple(t.shape)} fin={bool(torch.isfinite(t).all())} max={float(t.abs().max()) if t.numel() else 0:.3g}"
                        for t in ts if isinstance(t, torch.Tensor))
        ins = " ".join(f"{tuple(t.shape)}" for t in a if isinstance(t, torch.Tensor))
        extra = f" mode={a[2]} iters={a[3]}" if name == "ot_loop" else ""
        print(f"{name}({ins}{extra}) -> {info}", flush=True)
        return out

    setattr(mod, name, inner)


for n in ("fit_pca", "pca_project", "ot_loop"):
    wrap(_optex, n)
wrap(texture, "recentre")
sd = texture_cases.state_dicts()
torch.manual_seed(0)
model = texture.OptimalTexture(size=64, iters=10, passes=2, hist_mode="chol", no_multires=True, state_dicts=sd)
model.pca = _optex.fit_pca
for i, d in enumerate(model.decoders):
    f = d.forward

    def dec(x, _f=f, _i=i):
        out = _f(x)
        torch.cuda.synchronize()
        print(f"decoder[{_i}] -> fin={bool(torch.isfinite(out).all())} max={float(out.abs().max()):.3g}", flush=True)
        return out

    model.decoders[i] = dec
out = model.forward(torch.rand(1, 3, 64, 64, device="cuda"), [torch.rand(1, 3, 64, 96, device="cuda")])
print("final finite", bool(torch.isfinite(out).all()))
