import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimaltextures_b200 as ob
torch.set_printoptions(linewidth=200, precision=4, sci_mode=False)
for n, c in ((1600, 32), (1600, 31), (20000, 64)):
    g2 = torch.Generator().manual_seed(c)
    f = torch.relu(torch.randn(1, n, 1, c, generator=g2)).cuda()
    s = torch.relu(1.5 * torch.randn(1, n - 16, 1, c, generator=g2) + 0.25).cuda()
    out = ob.ot_loop(f, s, "pca", 1)
    P, S, O = f.reshape(-1, c).double(), s.reshape(-1, c).double(), out.reshape(-1, c).double()
    Pc, Oc = P - P.mean(0), O - O.mean(0)
    T_est = torch.linalg.lstsq(Pc, Oc).solution.T            # out_c = Pc T^T
    eye = torch.eye(c, device="cuda", dtype=torch.float64)
    ct = Pc.T @ Pc / n + eye
    cs = (S - S.mean(0)).T @ (S - S.mean(0)) / S.shape[0] + eye
    def sq(a):
        w, v = torch.linalg.eigh(a)
        return v @ torch.diag(w.sqrt()) @ v.T
    T_ref = sq(cs) @ torch.linalg.inv(sq(ct))
    d = (T_est - T_ref).abs()
    print(f"n={n} c={c}: |T_est - T_ref| max {float(d.max()):.2e}; per 8-row block {[round(float(d[i:i+8].max()),5) for i in range(0,c,8)]}; "
          f"per 8-col block {[round(float(d[:, i:i+8].max()),5) for i in range(0,c,8)]}")
    # what covariance would explain T_est?  T = Qs Qt^-1  ->  Qt = T^-1 Qs,  Sig_t = Qt Qt^T (if Qt symmetric)
    Qt = torch.linalg.inv(T_est) @ sq(cs)
    print("   asym of implied Qt", float((Qt - Qt.T).abs().max()), " implied Sig_t vs true: max diff", float((Qt @ Qt.T - ct).abs().max()),
          " diag diff", float((Qt @ Qt.T - ct).diag().abs().max()))
    # is it a convergence error?  residual of Z = T_est^-1... : Z_est = Qs^-1 T_est ; check Z Sig_t Z = I
    Z = torch.linalg.inv(sq(cs)) @ T_est
    print("   |Z Sig_t Z - I| max", float((Z @ ct @ Z - eye).abs().max()), " |Z - Z^T|", float((Z - Z.T).abs().max()))
