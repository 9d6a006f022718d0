"""Per-layer check of the encoder at a given size, B200 vs the CPU oracle, each layer fed the oracle's input and run
three times (determinism), under the toggles: PDL on/off, GEMM mode.  Test infrastructure."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from optimaltextures_b200 import _lib, vgg
from oracle import vgg_oracle

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sd = vgg_oracle.random_state_dict("encoder", depth)
enc = vgg.Encoder(depth, state_dict=sd)
x = torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(1))
wb = vgg_oracle.pairs(sd)
convs = vgg_oracle.encoder_convs(depth)
print("env chunk", os.environ.get("OPTEX_CONV_CHUNK_MB"), "no_a_tmem", os.environ.get("OPTEX_NO_A_TMEM"), flush=True)
for pdl in (1, 0):
    _lib.lib().optex_set_pdl(pdl)
    for mode in ("auto", "tf32", "fp32"):
        ob.set_gemm_mode(mode)
        cur = F.conv2d(x, wb[0][0], wb[0][1])
        msgs = []
        for i, ((pre, cin, cout, relu), (w, b)) in enumerate(zip(convs, wb[1:])):
            ref = vgg_oracle._layer(cur, pre, w, b, relu)
            if i == 0:
                outs = [enc.layers[0].run(x.cuda(), src_nchw=True) for _ in range(3)]
            else:
                inp = cur.permute(0, 2, 3, 1).contiguous().cuda()
                outs = [enc.layers[i].run(inp) for _ in range(3)]
            torch.cuda.synchronize()
            r = ref.permute(0, 2, 3, 1)
            errs = [float((o.cpu() - r).abs().max() / max(1.0, float(r.abs().max()))) for o in outs]
            same = torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
            bad = (outs[0] - outs[1]).abs() > 0
            where = ""
            if not same:
                rows = bad.reshape(-1, bad.shape[-1]).any(1).nonzero().flatten()
                where = f" differing rows {rows.numel()} first {rows[:4].tolist()} last {rows[-2:].tolist()}"
            msgs.append(f"L{i}({cin}->{cout},pre{pre},M={r.shape[1] * r.shape[2]}) err {max(errs):.1e} same {same}{where}")
            cur = ref
        print(f"pdl={pdl} gemm={mode}: " + " | ".join(msgs), flush=True)
