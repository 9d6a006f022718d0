"""Covariance modes at awkward channel counts / ragged pixel counts / large channel means, B200 vs CPU oracle
(single OT step with an injected rotation, and hist_match without).  Test infrastructure."""
import sys

import torch

sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from oracle import ot_oracle, rotation as rot_oracle

torch.manual_seed(0)
for c in (49, 41, 23, 64, 86, 33):
    for (hp, wp, hs, ws) in ((64, 64, 64, 96), (64, 64, 64, 64), (10, 12, 9, 14), (25, 40, 30, 50)):
        for kind in ("relu", "offset"):
            t = torch.relu(torch.randn(1, hp, wp, c))
            s = torch.relu(1.5 * torch.randn(1, hs, ws, c) + 0.25)
            if kind == "offset":    # PCA-projected features: the leading component carries the (scalar-centred) mean
                t[..., 0] += 40.0
                s[..., 0] += 38.0
                t[..., 1] -= 9.0
            r = torch.from_numpy(rot_oracle.haar_rotation_qr(c, 1)).float()
            msg = []
            for mode in ("chol", "pca", "sym"):
                ref = ot_oracle.ot_step(t, s, r, mode)
                out = ob.optimal_transport(t.cuda(), s.cuda(), mode, rotation=r.cuda()).cpu()
                e1 = float((out - ref).abs().max() / ref.abs().max())
                ref2 = ot_oracle.hist_match_nhwc(t, s, mode).contiguous()
                out2 = ob.hist_match(t.cuda(), s.cuda(), mode).cpu()
                e2 = float((out2 - ref2).abs().max() / ref2.abs().max())
                msg.append(f"{mode} step {e1:.1e} hm {e2:.1e}")
            print(f"c={c} n_p={hp * wp} n_s={hs * ws} {kind}: " + " | ".join(msg), flush=True)
