"""Numerical feasibility of moving the backward-preparation algebra onto warp-level MMAs.

CPU-only study (NumPy): realistic filtered covariances from the oracle's Kalman filter on a C2-shaped problem,
then per frame the algebra of kalman_backprep_rows2_kernel

    Wt = A S,  Pp = Wt A' + Q,  Lp = chol(Pp),  V = Lp^-1 Wt,  Sigma = S - V'V,  Ls = chol(Sigma),  G' = Lp^-T V

evaluated with the matrix products done in different arithmetics, block size 8 as an MMA implementation would
(diagonal blocks factored / inverted in scalar fp32, every off-diagonal update a small GEMM):

    f32      operands and accumulation in float32 (what the kernels do today)
    tf32     operands rounded to TF32 (10-bit mantissa), fp32 accumulation     (mma.sync .tf32, one term)
    tf32x3   3-term split a = hi + lo: hi*hi + hi*lo + lo*hi, fp32 accumulation  (fp32-grade emulation)
    f64acc   float32 operands, products and accumulation in float64, result rounded to float32 (DMMA)

Errors are against the same algebra in float64 and reported relative to the largest entry of each output.
Usage: python tools/numerics_mma_study.py > profiles/rNN_mma_numerics.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from keypoint_moseq_b200.synth import sample_dataset  # noqa: E402

BS = 8


def tf32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)


def mm(a, b, mode):
    """a @ b with a, b float32 arrays in the arithmetic `mode`; returns float32."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    if mode == "f32":
        return a @ b
    if mode == "tf32":
        return tf32(a) @ tf32(b)
    if mode == "tf32x3":
        ah, bh = tf32(a), tf32(b)
        al, bl = tf32(a - ah), tf32(b - bh)
        return (al @ bh + ah @ bl) + ah @ bh
    if mode == "f64acc":
        return (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
    raise ValueError(mode)


def chol_blocked(M, mode):
    """Right-looking blocked Cholesky (block 8): diagonal blocks in scalar fp32, panel = block times the
    explicit inverse of the diagonal factor, trailing update as GEMMs in `mode`."""
    n = M.shape[0]
    A = np.array(M, dtype=np.float32)
    L = np.zeros_like(A)
    for k0 in range(0, n, BS):
        k1 = min(k0 + BS, n)
        Lkk = np.linalg.cholesky(A[k0:k1, k0:k1].astype(np.float32)).astype(np.float32)
        L[k0:k1, k0:k1] = Lkk
        if k1 < n:
            inv = np.linalg.inv(Lkk.astype(np.float32)).astype(np.float32)
            L[k1:, k0:k1] = mm(A[k1:, k0:k1], inv.T, mode)
            A[k1:, k1:] = A[k1:, k1:] - mm(L[k1:, k0:k1], L[k1:, k0:k1].T, mode)
    return L


def trsm_lower(L, B, mode, transpose=False):
    """Solve L X = B (or L' X = B) by blocks: diagonal blocks through their explicit fp32 inverse, updates as GEMMs."""
    n = L.shape[0]
    X = np.array(B, dtype=np.float32)
    blocks = list(range(0, n, BS))
    order = blocks if not transpose else blocks[::-1]
    for k0 in order:
        k1 = min(k0 + BS, n)
        inv = np.linalg.inv(L[k0:k1, k0:k1].astype(np.float32)).astype(np.float32)
        X[k0:k1] = mm(inv.T if transpose else inv, X[k0:k1], mode)
        if not transpose and k1 < n:
            X[k1:] = X[k1:] - mm(L[k1:, k0:k1], X[k0:k1], mode)
        if transpose and k0 > 0:
            X[:k0] = X[:k0] - mm(L[k0:k1, :k0].T, X[k0:k1], mode)
    return X


def backprep(S, A, Q, mode):
    if mode == "f64":
        Wt = A @ S
        Pp = Wt @ A.T + Q
        Lp = np.linalg.cholesky(Pp)
        V = np.linalg.solve(Lp, Wt)
        Sig = S - V.T @ V
        Sig = 0.5 * (Sig + Sig.T)
        Ls = np.linalg.cholesky(Sig)
        GT = np.linalg.solve(Lp.T, V)
        return Sig, Ls, GT
    S32, A32, Q32 = (np.asarray(x, np.float32) for x in (S, A, Q))
    Wt = mm(A32, S32, mode)
    Pp = mm(Wt, A32.T, mode) + Q32
    Pp = 0.5 * (Pp + Pp.T)
    Lp = chol_blocked(Pp, mode)
    V = trsm_lower(Lp, Wt, mode)
    Sig = S32 - mm(V.T, V, mode)
    Sig = 0.5 * (Sig + Sig.T)
    Ls = chol_blocked(Sig, mode)
    GT = trsm_lower(Lp, V, mode, transpose=True)
    return Sig, Ls, GT


def main():
    cfg = dict(recordings=2, frames=400, k=12, D=2, d=10, L=3, K=100)
    data, _, model = sample_dataset(**cfg, seed=5, seg_length=400)
    Y, mask = np.asarray(data["Y"], np.float64), np.asarray(data["mask"], np.float64)
    st, pr = model["states"], model["params"]
    N, T, k, D = Y.shape
    d, L = cfg["d"], cfg["L"]
    n = d * L
    Ct = orc.lifted_obs_matrix(pr["Cd"], k, D)
    C = np.zeros((k * D, n))
    C[:, n - d:] = Ct[:, :-1]
    ys = orc.rotate(Y - st["v"][:, :, None, :], -st["h"]).reshape(N, T, k * D)[:, L - 1:]
    Rs = np.repeat(st["s"] * pr["sigmasq"], D, axis=-1)[:, L - 1:]
    A, B, Qa = orc.ar_to_lds(pr["Ab"], pr["Q"], 1e-3)
    fm, fS = orc.kalman_filter(ys, mask[:, L - 1:], st["z"], np.zeros(n), orc.X_PRIOR_VAR * np.eye(n), A, B, Qa, C,
                               Ct[:, -1], Rs)
    rng = np.random.default_rng(0)
    steady = [(int(rng.integers(N)), int(t)) for t in rng.integers(5, fS.shape[1] - 1, size=150)]
    start = [(nn, t) for nn in range(N) for t in range(4)]        # right after the 10 I prior: widest dynamic range
    for label, frames in (("steady-state frames", steady), ("first four frames of each chain", start)):
        study(label, frames, fS, st, A, Qa, n)


def study(label, frames, fS, st, A, Qa, n):
    modes = ["f32", "tf32", "tf32x3", "f64acc"]
    errs = {m: {"Sigma": [], "Ls": [], "G": []} for m in modes}
    failed = {m: 0 for m in modes}
    conds = []
    for nn, t in frames:
        S, j = fS[nn, t], st["z"][nn, t]
        ref = backprep(S, A[j], Qa[j], "f64")
        conds.append(np.linalg.cond(ref[0]))
        for m in modes:
            try:
                out = backprep(S, A[j], Qa[j], m)
            except np.linalg.LinAlgError:
                failed[m] += 1
                continue
            for name, a, b in zip(("Sigma", "Ls", "G"), out, ref):
                errs[m][name].append(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())
    print(f"# backward-preparation algebra on {len(frames)} {label} of a C2-shaped problem (n = {n}, block {BS}); "
          f"cond(Sigma): median {np.median(conds):.1e}, max {np.max(conds):.1e}")
    print("# error = max |variant - float64| / max |float64| per output; columns: median / 90th percentile / max over frames")
    print(f"{'mode':8s} {'Sigma':>32s} {'Ls = chol(Sigma)':>32s} {'G (gain)':>32s}  not-SPD")
    for m in modes:
        cells = []
        for name in ("Sigma", "Ls", "G"):
            e = np.array(errs[m][name]) if errs[m][name] else np.array([np.nan])
            cells.append(f"{np.median(e):9.1e} /{np.percentile(e, 90):9.1e} /{e.max():9.1e}")
        print(f"{m:8s} {cells[0]:>32s} {cells[1]:>32s} {cells[2]:>32s}  {failed[m]}")
    print()


if __name__ == "__main__":
    main()
