"""CPU checks of the time-parallel schemes the CUDA kernels implement, restated in
NumPy against the oracle's sequential recursions: they hold for any sizes, so the kernels' parity at full
size rests on these properties plus the small-size comparisons of test_gpu_parity.py."""
import numpy as np

import oracle as orc
from oracle.kpms_oracle import _categorical


def _random_hmm(rng, K, T, sticky=0.9, sharp=3.0):
    pi = rng.dirichlet(np.full(K, 0.3), size=K) * (1 - sticky) + sticky * np.eye(K)
    pi /= pi.sum(1, keepdims=True)
    ll = rng.standard_normal((1, T, K)) * sharp
    return pi, ll


def _backward_map(filt_t, pi, z_next, u_t):
    p = filt_t * pi[:, z_next]
    return int(_categorical(p / p.sum(), u_t))


def test_backward_sampling_by_merging_chunks_with_repair_is_exact():
    """With the uniforms fixed backward sampling is a deterministic map z_{t+1} -> z_t.  Chunks started
    from a guess W steps above their range, plus a top-down repair pass that re-walks from the true label
    until the new path meets the stored one, reproduce the sequential sampler whatever W is
    (hmm_backward_walk_kernel)."""
    rng = np.random.default_rng(0)
    for trial, (W, sticky) in enumerate([(0, 0.5), (3, 0.9), (8, 0.98), (20, 0.7)]):
        K, T, Lc = 6, 240, 40
        pi, ll = _random_hmm(rng, K, T, sticky)
        u = rng.uniform(size=(1, T))
        z_ref, _ = orc.sample_hmm_stateseq(pi, ll, np.ones((1, T)), u)
        _, filt = orc.hmm_filter(pi, ll)
        filt, u1, z_ref = filt[0], u[0], z_ref[0]
        z = np.full(T, -1)
        C = T // Lc
        zwarm = np.full(C, -1)
        for c in range(C):                                        # every chunk independently
            begin, end = c * Lc, (c + 1) * Lc
            top = end == T
            t0 = T - 1 if top else min(end - 1 + max(W, 1), T - 1)
            zc = int(_categorical(filt[t0], u1[t0]))           # draw from the filtered marginal alone
            if t0 < end:
                z[t0] = zc
            elif t0 == end:
                zwarm[c] = zc
            for t in range(t0 - 1, begin - 1, -1):
                zc = _backward_map(filt[t], pi, zc, u1[t])
                if t < end:
                    z[t] = zc
                elif t == end:
                    zwarm[c] = zc
        mismatches = rewalked = 0
        for c in range(C - 2, -1, -1):                            # repair, top down
            begin, end = c * Lc, (c + 1) * Lc
            zc = z[end]
            if zwarm[c] == zc:
                continue
            mismatches += 1
            for t in range(end - 1, begin - 1, -1):
                zt = _backward_map(filt[t], pi, zc, u1[t])
                if zt == z[t]:
                    break                                          # met the stored path
                z[t], zc = zt, zt
                rewalked += 1
        assert np.array_equal(z, z_ref), (trial, mismatches, rewalked)
        if W == 0:
            assert mismatches > 0                                  # the repair pass was exercised


def test_stay_test_is_the_inverse_cdf_decision():
    """The walker accepts "the label stays j" iff c_{j-1} < r <= c_j; that is exactly the event
    #{i : c_i < r} == j of the inverse-CDF rule."""
    rng = np.random.default_rng(1)
    K = 9
    for _ in range(2000):
        filt = rng.dirichlet(np.full(K, 0.2))
        col = rng.dirichlet(np.full(K, 0.2))
        j = int(rng.integers(K))
        u = rng.uniform()
        c = np.cumsum(filt * col)
        r = c[-1] * (1.0 - u)
        label = int((c < r).sum())
        lo = c[j - 1] if j > 0 else 0.0
        stay = (lo < r) and not (c[j] < r)
        assert stay == (label == j)


def test_filter_chunks_forget_their_start_and_the_boundary_check_sees_it():
    """A filter chunk started W steps early from the uniform prior approaches the sequential filter as W
    grows.  The filter step is non-expansive in Hilbert's projective metric
    d(p, q) = max log(p/q) - min log(p/q), so a boundary state that agrees COMPONENT BY COMPONENT
    (relative tolerance, boundary_check_kernel<COMPONENTWISE>) bounds every later relative error - the
    sup-norm difference alone does not (it can grow when a small component is later favoured)."""
    rng = np.random.default_rng(2)
    K, T = 8, 400
    pi, ll = _random_hmm(rng, K, T, sticky=0.9, sharp=1.0)
    _, filt = orc.hmm_filter(pi, ll)
    filt = filt[0]
    begin = 200

    def run(start, p0, stop):
        pred, out = p0, {}
        for t in range(start, stop):
            q = pred * np.exp(ll[0, t] - ll[0, t].max())
            f = q / q.sum()
            out[t] = f
            pred = f @ pi
        return out

    def hilbert(p, q):
        r = np.log(p) - np.log(q)
        return r.max() - r.min()

    errs = []
    for W in (4, 16, 64):
        out = run(begin - W, np.full(K, 1.0 / K), begin + 100)
        d0 = hilbert(out[begin - 1], filt[begin - 1])
        rel0 = np.abs(out[begin - 1] / filt[begin - 1] - 1).max()
        for t in range(begin, begin + 100):
            assert hilbert(out[t], filt[t]) <= d0 * (1 + 1e-9) + 1e-13          # never expands
            assert np.abs(out[t] / filt[t] - 1).max() <= 2.0 * rel0 * (1 + 1e-6) + 1e-12
        errs.append(np.abs(out[begin - 1] - filt[begin - 1]).max())
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-4
    # the sup norm is NOT monotone: a start that is close in sup norm but wrong in a small component
    p = filt[begin - 1].copy()
    j = int(np.argmin(p))
    q = p.copy()
    q[j] *= 50.0
    q /= q.sum()
    ll2 = ll.copy()
    ll2[0, begin:begin + 3, j] += 6.0                                          # the small state is favoured next
    pa, pb = p @ pi, q @ pi
    grow = []
    for t in range(begin, begin + 3):
        wa = pa * np.exp(ll2[0, t] - ll2[0, t].max()); wa /= wa.sum()
        wb = pb * np.exp(ll2[0, t] - ll2[0, t].max()); wb /= wb.sum()
        grow.append(np.abs(wa - wb).max())
        pa, pb = wa @ pi, wb @ pi
    assert max(grow) > np.abs(p - q).max()                                     # amplified in sup norm ...
    assert hilbert(wa, wb) <= hilbert(p, q) * (1 + 1e-9)                        # ... but not in the projective metric


def test_padded_tail_chunks_start_from_powers_of_pi():
    """Masked frames carry no likelihood, so p_{t+TL} = (pi^TL)' p_t exactly: the padded tail is cut into
    TL-step chunks whose starting predictions come from the precomputed power (hmm_tail_starts_kernel)."""
    rng = np.random.default_rng(3)
    K, TL = 7, 16
    pi, _ = _random_hmm(rng, K, 1)
    p = rng.dirichlet(np.ones(K))
    seq = p.copy()
    for _ in range(TL):
        seq = seq @ pi
    P = np.linalg.matrix_power(pi, TL)
    assert np.allclose(p @ P, seq, rtol=1e-13, atol=1e-15)
    sq = pi.copy()
    for _ in range(4):                                             # repeated squaring as the kernel does
        sq = sq @ sq
    assert np.allclose(sq, P, rtol=1e-12)


def test_affine_recursion_chunks_and_centroid_scans_are_associative_maps():
    """xi_t = G_t xi_{t+1} + h_t composes as affine maps, so chunks can start anywhere once their
    boundary value is right, and scans over (G, h) pairs give the same path as the serial recursion
    (kalman_affine_kernel, location_ffbs_kernel)."""
    rng = np.random.default_rng(4)
    n, T = 5, 90
    G = rng.standard_normal((T, n, n)) * 0.3
    h = rng.standard_normal((T, n))
    xi = np.empty((T, n))
    xi[-1] = h[-1]
    for t in range(T - 2, -1, -1):
        xi[t] = G[t] @ xi[t + 1] + h[t]
    # compose maps of a chunk, apply once
    lo, hi = 10, 30                                                # xi[lo] from xi[hi]
    A, b = np.eye(n), np.zeros(n)
    for t in range(hi - 1, lo - 1, -1):                            # x -> G_t x + h_t applied after the maps above it
        A, b = G[t] @ A, G[t] @ b + h[t]
    assert np.allclose(A @ xi[hi] + b, xi[lo], rtol=1e-10, atol=1e-12)
    # a chunk started W steps above from zero converges when the maps contract
    W = 40
    x = np.zeros(n)
    for t in range(hi + W - 1, hi - 1, -1):
        x = G[t] @ x + h[t]
    assert np.abs(x - xi[hi]).max() < 1e-5 * (1 + np.abs(xi[hi]).max())


def test_refinement_passes_converge_and_the_boundary_check_certifies_them():
    """Planned fallback for chains that do not forget within the warm-up: every chunk
    restarts from the end state its predecessor produced in the previous pass.  The boundary discrepancy
    contracts pass after pass, and once every chunk's start agrees component by component with its
    predecessor's new end state the whole filter equals the sequential one to that tolerance."""
    rng = np.random.default_rng(5)
    K, T, Lc, W = 6, 360, 60, 4
    pi, ll = _random_hmm(rng, K, T, sticky=0.97, sharp=0.6)          # sticky, weakly informative: slow forgetting
    _, filt = orc.hmm_filter(pi, ll)
    filt = filt[0]
    C = T // Lc

    def run(start, pred, stop, out):
        for t in range(start, stop):
            q = pred * np.exp(ll[0, t] - ll[0, t].max())
            f = q / q.sum()
            if t >= 0:
                out[t] = f
            pred = f @ pi
        return pred                                                   # prediction handed to the next chunk

    est = np.zeros((T, K))
    ends = [None] * C
    for c in range(C):                                                # pass 0: warm-up from the uniform prior
        tmp = {}
        start = max(c * Lc - W, 0)
        ends[c] = run(start, np.full(K, 1.0 / K), (c + 1) * Lc, tmp)
        for t in range(c * Lc, (c + 1) * Lc):
            est[t] = tmp[t]
    err0 = np.abs(est / filt - 1).max()
    assert err0 > 1e-6                                                # the short warm-up is not enough here
    passes = 0
    while True:
        passes += 1
        starts = [np.full(K, 1.0 / K)] + [ends[c - 1] for c in range(1, C)]
        new_ends = list(ends)
        for c in range(1, C):
            tmp = {}
            new_ends[c] = run(c * Lc, starts[c], (c + 1) * Lc, tmp)
            for t in range(c * Lc, (c + 1) * Lc):
                est[t] = tmp[t]
        # the check: the start every chunk used against its predecessor's end state of THIS pass
        disc = max(np.abs(starts[c] / new_ends[c - 1] - 1).max() for c in range(1, C))
        ends = new_ends
        if disc <= 1e-12 or passes > C:
            break
    assert passes <= C                                                # at worst one chunk per pass (sequential)
    assert np.abs(est / filt - 1).max() < 1e-10                       # certified by the boundary check alone
