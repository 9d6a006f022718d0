"""CPU simulation (numpy, FP64) of the PCA eigensolver's sweep count: the kernel's one-sided Jacobi on the rows of
W = G (pca.cu) against the same iteration on the rows of the Cholesky factor R of G = R^T R (Drmac-Veselic style
preconditioning: R R^T is one LR step closer to diagonal than G, the singular values are sigma instead of
lambda = sigma^2, and the eigenvectors are the normalised rows of the final W - no V to carry).  Same round-robin
order, same JTOL / early-exit rule as the kernel.  Also a BLOCKED form (b rows per block, a pair of blocks solved
exactly per step: c / b - 1 rounds per sweep instead of c - 1).  `--blocked-order`: the pairing ORDER of round 2's
pca_jacobi_blk_kernel against the round-robin order (sweep counts, lambda_i = |w_i| without V).
Usage: python scripts/jacobi_sim.py [--blocked-order] [c] [n ...]"""
import sys

import numpy as np

JTOL, JEXIT = 1e-12, 1e-6


def schedule(c):
    ne = c + (c & 1)
    n1, m = ne - 1, ne // 2
    for r in range(n1):
        pi = np.arange(m)
        a = np.where(pi == 0, n1, (r + pi) % n1)
        b = np.where(pi == 0, r, (r - pi + n1) % n1)
        ok = (a < c) & (b < c)
        yield a[ok], b[ok]


def jacobi_rows(W, V=None, max_sweeps=40, floor_rel=1e-13):
    c = W.shape[0]
    tr = np.sqrt((W * W).sum()) if V is None else np.trace(W)
    floor_abs = (floor_rel * tr) ** 2
    rounds = list(schedule(c))
    for sweep in range(max_sweeps):
        rot, mx = 0, 0.0
        for a, b in rounds:
            wa, wb = W[a], W[b]
            al, be, ga = (wa * wa).sum(1), (wb * wb).sum(1), (wa * wb).sum(1)
            go = (ga * ga > JTOL * JTOL * al * be) & (np.abs(ga) > floor_abs)
            if not go.any():
                continue
            a2, b2, al, be, ga = a[go], b[go], al[go], be[go], ga[go]
            zeta = (be - al) / (2 * ga)
            t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
            cs = 1 / np.sqrt(1 + t * t)
            sn = cs * t
            wa, wb = W[a2], W[b2]
            W[a2], W[b2] = cs[:, None] * wa - sn[:, None] * wb, sn[:, None] * wa + cs[:, None] * wb
            if V is not None:
                va, vb = V[a2], V[b2]
                V[a2], V[b2] = cs[:, None] * va - sn[:, None] * vb, sn[:, None] * va + cs[:, None] * vb
            rot += int(go.sum())
            mx = max(mx, float((ga * ga / (al * be)).max()))
        if rot == 0 or mx <= JEXIT * JEXIT:
            return sweep + 1
    return max_sweeps


def block_jacobi_rows(W, b, max_sweeps=30, floor_rel=1e-13):
    """Block one-sided Jacobi: the 2b rows of a block pair are orthogonalised exactly (eigh of their Gram matrix, the
    eigenvector matrix permuted / signed towards the identity, as convergence of block methods needs)."""
    c = W.shape[0]
    tr = np.sqrt((W * W).sum())
    rounds = list(schedule(c // b))
    for sweep in range(max_sweeps):
        mx = 0.0
        for A, B in rounds:
            for a, bb in zip(A, B):
                idx = np.r_[a * b:(a + 1) * b, bb * b:(bb + 1) * b]
                X = W[idx]
                S = X @ X.T
                d = np.sqrt(np.diag(S))
                live = d > floor_rel * tr                      # null rows are rounding noise: not part of the vote
                if live.sum() > 1:
                    off = np.abs(S[np.ix_(live, live)] / np.outer(d[live], d[live]))
                    np.fill_diagonal(off, 0)
                    mx = max(mx, float(off.max()))
                _, U = np.linalg.eigh(S)
                M, perm, used = np.abs(U), -np.ones(2 * b, int), set()
                for i in np.argsort(-M.max(1)):
                    j = [j for j in np.argsort(-M[i]) if j not in used][0]
                    perm[i] = j
                    used.add(j)
                U = U[:, perm]
                U = U * np.where(np.diag(U) < 0, -1.0, 1.0)[None, :]
                W[idx] = U.T @ X
        if mx <= JEXIT:
            return sweep + 1
    return max_sweeps


def features(n, c, seed):
    g = np.random.default_rng(seed)
    mix = g.standard_normal((c, c)) * (2.0 / np.sqrt(c))
    return np.maximum(g.standard_normal((n, c)) @ mix + 0.3, 0).astype(np.float32)


_args = [v for v in sys.argv[1:] if not v.startswith("--")]
c = int(_args[0]) if _args else 128
ns = [int(v) for v in _args[1:]] or [c * 3 // 4, 3 * c, 8 * c]
def blocked_order_rows(W, bsz, max_sweeps=40, floor_rel=1e-13):
    """The ORDER of pca.cu's pca_jacobi_blk_kernel (round 2): blocks of `bsz` rows paired round-robin; a block pair
    rotates its bsz x bsz cross pairs in bsz sub-rounds of bsz disjoint pairs; global round 0 of every sweep runs the
    full tournament of the 2 bsz rows instead (the pairs inside each block).  Same rotations and stopping rule as
    jacobi_rows - only the order differs (and with it the number of grid-wide barriers: c / bsz - 1 per sweep)."""
    c = W.shape[0]
    nb = c // bsz
    floor_abs = (floor_rel * np.trace(W)) ** 2
    brounds = list(schedule(nb))
    inner = list(schedule(2 * bsz))

    def rotate(a, b, st):
        wa, wb = W[a], W[b]
        al, be, ga = (wa * wa).sum(1), (wb * wb).sum(1), (wa * wb).sum(1)
        go = (ga * ga > JTOL * JTOL * al * be) & (np.abs(ga) > floor_abs)
        if not go.any():
            return
        a2, b2, al, be, ga = a[go], b[go], al[go], be[go], ga[go]
        zeta = (be - al) / (2 * ga)
        t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
        cs = 1 / np.sqrt(1 + t * t)
        sn = cs * t
        wa, wb = W[a2], W[b2]
        W[a2], W[b2] = cs[:, None] * wa - sn[:, None] * wb, sn[:, None] * wa + cs[:, None] * wb
        st[0] += int(go.sum())
        st[1] = max(st[1], float((ga * ga / (al * be)).max()))

    for sweep in range(max_sweeps):
        st = [0, 0.0]
        for gr, (A, B) in enumerate(brounds):
            if gr == 0:
                for ia, ib in inner:
                    ra = np.where(ia < bsz, A[:, None] * bsz + ia, B[:, None] * bsz + ia - bsz).ravel()
                    rb = np.where(ib < bsz, A[:, None] * bsz + ib, B[:, None] * bsz + ib - bsz).ravel()
                    rotate(ra, rb, st)
            else:
                i = np.arange(bsz)
                for sr in range(bsz):
                    rotate((A[:, None] * bsz + i).ravel(), (B[:, None] * bsz + (i + sr) % bsz).ravel(), st)
        if st[0] == 0 or st[1] <= JEXIT * JEXIT:
            return sweep + 1
    return max_sweeps


if "--blocked-order" in sys.argv:
    for n in ns:
        x = features(n, c, 0).astype(np.float64)
        xc = x - x.mean()
        G = xc.T @ xc
        lam = np.linalg.eigvalsh(G)[::-1]
        row = [f"c={c} n={n}: round-robin {jacobi_rows(G.copy(), np.eye(c))} sweeps"]
        for bsz in (4, 8, 16):
            if c % (2 * bsz):
                continue
            W = G.copy()
            sw = blocked_order_rows(W, bsz)
            lw = np.sort(np.sqrt((W * W).sum(1)))[::-1]          # lambda_i = |w_i| (no V)
            row.append(f"blocks of {bsz}: {sw} sweeps, {c // bsz - 1} barriers / sweep, "
                       f"dlambda {np.abs(lw - lam).max() / lam[0]:.1e}")
        print(" | ".join(row), flush=True)
    sys.exit(0)


for n in ns:
    x = features(n, c, 0).astype(np.float64)
    xc = x - x.mean()
    G = xc.T @ xc
    lam_ref = np.linalg.eigvalsh(G)[::-1]
    # kernel's formulation
    W, V = G.copy(), np.eye(c)
    s0 = jacobi_rows(W, V)
    lam0 = np.sort((V * W).sum(1))[::-1]
    # Cholesky-preconditioned: rows of R, G + jitter = R^T R
    jit = 1e-13 * np.trace(G)
    R = np.linalg.cholesky(G + jit * np.eye(c)).T
    W1 = R.copy()
    s1 = jacobi_rows(W1, None)
    lam1 = np.sort((W1 * W1).sum(1))[::-1] - jit
    Q = W1 / np.linalg.norm(W1, axis=1, keepdims=True)
    order = np.argsort(-(W1 * W1).sum(1))
    k = int(min(n, c) * 0.8)
    Qk = Q[order[:k]]
    resid = np.abs(Qk @ G @ Qk.T - np.diag(np.diag(Qk @ G @ Qk.T))).max() / lam_ref[0]
    orth = np.abs(Qk @ Qk.T - np.eye(k)).max()
    sig = lambda l: np.sqrt(np.clip(l, 0, None))
    print(f"c={c} n={n}: rows of G: {s0} sweeps (dsigma {np.abs(sig(lam0) - sig(lam_ref)).max() / sig(lam_ref)[0]:.1e}) | "
          f"rows of chol(G): {s1} sweeps (dsigma {np.abs(sig(lam1) - sig(lam_ref)).max() / sig(lam_ref)[0]:.1e}, "
          f"top-{k} offdiag {resid:.1e}, orth {orth:.1e})", flush=True)


def pivoted_cholesky(G):
    """G[p][:, p] = R^T R with the diagonal of R non-increasing (complete pivoting); returns R, p."""
    A = G.copy()
    c = A.shape[0]
    p = np.arange(c)
    R = np.zeros_like(A)
    for j in range(c):
        i = j + int(np.argmax(np.diag(A)[j:]))
        if i != j:
            A[[j, i]] = A[[i, j]]
            A[:, [j, i]] = A[:, [i, j]]
            R[:, [j, i]] = R[:, [i, j]]
            p[[j, i]] = p[[i, j]]
        d = A[j, j]
        if d <= 0:
            break
        R[j, j] = np.sqrt(d)
        R[j, j + 1:] = A[j, j + 1:] / R[j, j]
        A[j + 1:, j + 1:] -= np.outer(R[j, j + 1:], R[j, j + 1:])
    return R, p


if "--pivot" in sys.argv or True:
    for n in ns:
        x = features(n, c, 0).astype(np.float64)
        xc = x - x.mean()
        G = xc.T @ xc
        jit = 1e-13 * np.trace(G)
        Gj = G + jit * np.eye(c)
        # (1) diagonal-sorted, (2) completely pivoted
        p1 = np.argsort(-np.diag(Gj))
        R1 = np.linalg.cholesky(Gj[p1][:, p1]).T
        R2, p2 = pivoted_cholesky(Gj)
        print(f"c={c} n={n}: chol of diag-sorted G: {jacobi_rows(R1.copy())} sweeps | pivoted chol: "
              f"{jacobi_rows(R2.copy())} sweeps | rows of R^T (= L, other orientation): "
              f"{jacobi_rows(np.linalg.cholesky(Gj).copy())} sweeps", flush=True)


for b in (8, 16):
    if c % (2 * b):
        continue
    for n in ns:
        x = features(n, c, 0).astype(np.float64)
        xc = x - x.mean()
        G = xc.T @ xc
        R = np.linalg.cholesky(G + 1e-13 * np.trace(G) * np.eye(c)).T
        print(f"c={c} n={n} blocked b={b} ({c // b - 1} rounds per sweep): rows of G {block_jacobi_rows(G.copy(), b)} "
              f"sweeps | rows of chol(G) {block_jacobi_rows(R.copy(), b)} sweeps", flush=True)
