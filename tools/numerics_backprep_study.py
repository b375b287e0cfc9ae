"""Float32 accuracy of two factorisations of the backward-sampling step of the Kalman FFBS.

CPU-only study (NumPy).  Realistic filtered moments (m_t, S_t) come from the oracle's float64 Kalman filter on a
C2-shaped chain; per frame the conditional xi_t | xi_{t+1} = N(G xi_{t+1} + h0, Sigma) is evaluated

  alg0  one n-dimensional update (what kalman_backprep_rows2_kernel did in round 1):
        Wt = Aaug S, Pp = Wt Aaug' + Qaug, Lp = chol(Pp), V = Lp^-1 Wt, Sigma = S - V'V, G' = Lp^-T V
  alg1  two-stage update that uses the companion structure xi_{t+1} = [c + e1; A xi_t + b + e2], xi_t = [a; c]:
        stage 1 conditions on the (n-d) shifted coordinates (noise eps I) in closed form through Z = (S_cc + eps I)^-1,
        stage 2 on the d new coordinates (noise Q') through a d x d factorisation.

both in float32 (every product and factorisation on float32 operands) against float64.  Reported: error of G, of
chol(Sigma) and of the offset h0 = m - G (Aaug m + b), each relative to the largest entry of the float64 result,
as median / 99th percentile / max over frames.
Usage: python tools/numerics_backprep_study.py > profiles/r02_backprep_numerics.txt
"""
import os
import sys

import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from keypoint_moseq_b200.synth import sample_dataset  # noqa: E402

F = np.float32


def filtered_moments(frames=700, seed=5, d=10, L=3, k=12, D=2, K=100):
    data, _, model = sample_dataset(recordings=1, frames=frames, k=k, D=D, d=d, L=L, K=K, seed=seed, seg_length=frames)
    st, pr = model["states"], model["params"]
    N, T = data["Y"].shape[:2]
    n = d * L
    Ct = orc.lifted_obs_matrix(pr["Cd"], k, D)
    C = np.zeros((k * D, n))
    C[:, n - d:] = Ct[:, :-1]
    ys = orc.rotate(data["Y"] - st["v"][:, :, None, :], -st["h"]).reshape(N, T, k * D)[:, L - 1:]
    Rs = np.repeat(st["s"] * pr["sigmasq"], D, axis=-1)[:, L - 1:]
    A, B, Qa = orc.ar_to_lds(pr["Ab"], pr["Q"], 1e-3)
    fm, fS = orc.kalman_filter(ys, data["mask"][:, L - 1:], st["z"], np.zeros(n), orc.X_PRIOR_VAR * np.eye(n),
                               A, B, Qa, C, Ct[:, -1], Rs)
    return fm[0], fS[0], st["z"][0], pr["Ab"], pr["Q"]


def truth(m, S, Ab, Q, eps, jit):
    d = Ab.shape[0]
    n = Ab.shape[1] - 1
    NO = n - d
    Aaug = np.zeros((n, n))
    Aaug[:NO, d:] = np.eye(NO)
    Aaug[NO:] = Ab[:, :n]
    baug = np.concatenate([np.zeros(NO), Ab[:, n]])
    Qaug = np.zeros((n, n))
    Qaug[:NO, :NO] = np.eye(NO) * eps
    Qaug[NO:, NO:] = Q
    Qaug += np.eye(n) * jit
    Pp = Aaug @ S @ Aaug.T + Qaug
    G = np.linalg.solve(Pp, Aaug @ S).T
    Sig = S - G @ Pp @ G.T
    Sig = 0.5 * (Sig + Sig.T)
    return G, np.linalg.cholesky(Sig), m - G @ (Aaug @ m + baug)


def alg0(m, S, Ab, Q, eps, jit):
    m, S, Ab, Q = F(m), F(S), F(Ab), F(Q)
    d = Ab.shape[0]
    n = Ab.shape[1] - 1
    NO = n - d
    A = Ab[:, :n]
    Wt = np.concatenate([S[d:], A @ S], axis=0)
    Pp = np.concatenate([Wt[:, d:], Wt @ A.T], axis=1)
    Pp[:NO, :NO] += np.eye(NO, dtype=F) * F(eps + jit)
    Pp[NO:, NO:] += Q + np.eye(d, dtype=F) * F(jit)
    Lp = np.linalg.cholesky(Pp)
    V = sl.solve_triangular(Lp, Wt, lower=True).astype(F)
    Sig = S - V.T @ V
    Ls = np.linalg.cholesky(Sig)
    GT = sl.solve_triangular(Lp, V, lower=True, trans="T").astype(F)
    mp = np.concatenate([m[d:], A @ m + Ab[:, n]])
    return GT.T, Ls, m - GT.T @ mp


def gauss_jordan(M, NO):
    """In-place Gauss-Jordan on [B | R] (NO x (NO + r)), no pivoting (B is SPD): returns [B^-1 | B^-1 R], every
    operation rounded to float32 as the kernel does it (row scaling by the reciprocal pivot, one FMA per entry)."""
    M = np.array(M, dtype=F)
    for j in range(NO):
        ip = F(1.0) / M[j, j]
        row = M[j] * ip
        row[j] = ip
        f = M[:, j].copy()
        M -= np.outer(f, row).astype(F)
        M[:, j] = -f * ip
        M[j] = row
    return M


def alg1(m, S, Ab, Q, eps, jit, gj=False):
    m, S, Ab, Q = F(m), F(S), F(Ab), F(Q)
    d = Ab.shape[0]
    n = Ab.shape[1] - 1
    NO = n - d
    e1 = F(eps + jit)
    Aa, Ac = Ab[:, :d], Ab[:, d:n]
    Saa, Sac, Scc = S[:d, :d], S[:d, d:], S[d:, d:]
    B1 = Scc + np.eye(NO, dtype=F) * e1
    if gj:
        M = gauss_jordan(np.concatenate([B1, Sac.T], axis=1), NO)
        Z, T = M[:, :NO], M[:, NO:].T.copy()
    else:
        L1 = np.linalg.cholesky(B1)
        U = sl.solve_triangular(L1, np.eye(NO, dtype=F), lower=True).astype(F)      # L1^-1
        Z = U.T @ U
        T = Sac @ Z
    Fm = np.eye(NO, dtype=F) - e1 * Z
    S1aa = Saa - T @ Sac.T
    Wc = Aa @ T + Ac @ Fm                                                      # A K1
    Wa = Aa @ S1aa + e1 * (Ac @ T.T)
    B2 = Wa @ Aa.T + e1 * (Wc @ Ac.T) + Q + np.eye(d, dtype=F) * F(jit)
    B2 = F(0.5) * (B2 + B2.T)
    L2 = np.linalg.cholesky(B2)
    W = np.concatenate([Wa, e1 * Wc], axis=1)                                   # d x n
    V2 = sl.solve_triangular(L2, W, lower=True).astype(F).T                     # n x d
    K2 = sl.solve_triangular(L2, V2.T, lower=True, trans="T").astype(F).T       # n x d
    S1 = np.empty((n, n), dtype=F)
    S1[:d, :d] = S1aa
    S1[:d, d:] = e1 * T
    S1[d:, :d] = e1 * T.T
    S1[d:, d:] = e1 * Fm
    Sig = S1 - V2 @ V2.T
    Sig = F(0.5) * (Sig + Sig.T)
    Ls = np.linalg.cholesky(Sig)
    K1 = np.concatenate([T, Fm], axis=0)                                        # n x NO
    G = np.concatenate([K1 - K2 @ Wc, K2], axis=1)
    mp = np.concatenate([m[d:], Ab[:, :n] @ m + Ab[:, n]])
    return G, Ls, m - G @ mp


def main():
    eps, jit = orc.EPS_SHIFT, 1e-3
    for d, tag in ((10, "latent_dim 10 (C2)"), (4, "latent_dim 4 (C1)")):
        fm, fS, z, Ab, Q = filtered_moments(d=d, k=12 if d == 10 else 10)
        rows = {"alg0": [], "alg1": [], "alg1-gj": []}
        for t in range(fS.shape[0] - 1):
            ref = truth(fm[t], fS[t], Ab[z[t]], Q[z[t]], eps, jit)
            for name, fn in (("alg0", alg0), ("alg1", alg1), ("alg1-gj", lambda *a: alg1(*a, gj=True))):
                out = fn(fm[t], fS[t], Ab[z[t]], Q[z[t]], eps, jit)
                rows[name].append([np.abs(o - r).max() / np.abs(r).max() for o, r in zip(out, ref)])
        print(f"== {tag}: {fS.shape[0] - 1} frames, filtered covariance diag range "
              f"{np.diagonal(fS, axis1=1, axis2=2).min():.3g} .. {np.diagonal(fS, axis1=1, axis2=2).max():.3g}")
        for name, r in rows.items():
            r = np.array(r)
            for j, what in enumerate(("G", "chol(Sigma)", "h0")):
                print(f"  {name} {what:12s} median {np.median(r[:, j]):.2e}  p99 {np.percentile(r[:, j], 99):.2e}  max {r[:, j].max():.2e}"
                      f"  (first 5 frames max {r[:5, j].max():.2e})")


if __name__ == "__main__":
    main()
