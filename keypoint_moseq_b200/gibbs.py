"""One keypoint-SLDS Gibbs sweep on a B200: the drop-in for
`jax_moseq.models.keypoint_slds.resample_model`, which the reference calls once per
iteration as `model = resample_func(data, **model, **resample_options)`
(/root/reference/keypoint_moseq/fitting.py:25; bound at :245-248, :396-399, :510-513).

The function names, argument meaning and the returned dict mirror the upstream samplers;
every sampler is a thin host wrapper over the C-ABI in include/kpms_b200.h.  Buffers are
torch CUDA tensors; there is no CPU path.
"""
import numpy as np
import torch

from . import _lib
from .synth import center_embedding

__all__ = [
    "resample_model", "resample_discrete_stateseqs", "resample_continuous_stateseqs",
    "resample_scales", "resample_heading_location", "resample_ar_params",
    "resample_hdp_transitions", "resample_obs_variance", "sufficient_statistics",
    "marginal_log_likelihood", "stateseq_marginals", "lifted_obs_matrix", "seed_to_u64", "release_graphs",
    "advance_seed", "to_device_model", "to_device_data", "chunk_diagnostics",
]

_SCRATCH = {}
_SCRATCH_EPOCH = [0]


def _scratch(tag, nbytes, device):
    """Persistent byte workspace per (tag, device); grows monotonically."""
    key = (tag, str(device))
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _SCRATCH_EPOCH[0] += 1          # captured graphs hold the old pointer: they re-capture before their next replay
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf


def chunk_diagnostics(tag="kalman_ws", device="cuda"):
    """Diagnostics of the last time-chunked call that used scratch `tag` (synchronises):
    largest boundary discrepancy and number of chains re-run sequentially, forward and backward."""
    buf = _SCRATCH.get((tag, str(torch.device(device) if not isinstance(device, torch.device) else device)))
    if buf is None:
        for (t, _), b in _SCRATCH.items():
            if t == tag:
                buf = b
    if buf is None:
        return None
    raw = buf[:32].cpu().numpy()
    u = raw.view(np.uint32)
    f = raw.view(np.float32)
    if tag == "hmm_ws":     # backward sampler: exact merge-and-repair, counts instead of a tolerance
        out = {"forward_max_err": float(f[0]), "forward_rerun": int(u[1]),
               "backward_mismatches": int(u[2]), "backward_rewalked": int(u[3])}
        if int(u[4]) > 0:   # refinement passes enabled (KPMS_HMM_REFINE): chains flagged first / still flagged after
            out.update(forward_refine_passes=int(u[4]), forward_flagged_first=int(u[1]), forward_rerun=int(u[5]))
        return out
    return {"forward_max_err": float(f[0]), "forward_rerun": int(u[1]),
            "backward_max_err": float(f[2]), "backward_rerun": int(u[3])}


_STREAMS = {}


def _side_stream(dev, tag):
    key = (str(dev), tag)
    if key not in _STREAMS:
        _STREAMS[key] = torch.cuda.Stream(device=dev)
    return _STREAMS[key]


class _Stager:
    """Uploads host operands on a copy stream in the order they are registered; `get` makes the
    compute stream wait for that operand only.  Device operands pass through (converted if needed)."""

    def __init__(self, dev):
        self.dev, self.items, self.copy = dev, {}, None
        self.main = torch.cuda.current_stream(dev)

    def put(self, key, a, dtype):
        if a is None:
            self.items[key] = (None, None)
            return
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a))
        if a.is_cuda:
            self.items[key] = (a.to(device=self.dev, dtype=dtype).contiguous(), None)
            return
        if self.copy is None:
            self.copy = _side_stream(self.dev, "h2d")
        with torch.cuda.stream(self.copy):
            t = a.to(self.dev, non_blocking=True)
            if t.dtype != dtype:
                t = t.to(dtype)
            t = t.contiguous()
            ev = torch.cuda.Event()
            ev.record(self.copy)
        t.record_stream(self.main)
        self.items[key] = (t, ev)

    def get(self, key):
        t, ev = self.items[key]
        if ev is not None:
            self.main.wait_event(ev)
            self.items[key] = (t, None)
        return t


class _HostSink:
    """Copies finished states into caller-provided host tensors on a side stream."""

    def __init__(self, dev, host_out):
        self.out, self.done = host_out, set()
        if host_out is not None:
            self.main = torch.cuda.current_stream(dev)
            self.side = _side_stream(dev, "d2h")

    def emit(self, key, t):
        if self.out is None or key not in self.out or key in self.done:
            return
        self.done.add(key)
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            self.out[key].copy_(t, non_blocking=True)
        t.record_stream(self.side)

    def finish(self, states):
        if self.out is None:
            return
        for key, t in states.items():
            self.emit(key, t)


def _dev(a, dtype, device):
    """Tensor on `device` with `dtype`, contiguous; no copy when already so."""
    if a is None:
        return None
    if not isinstance(a, torch.Tensor):
        a = torch.as_tensor(np.asarray(a))
    return a.to(device=device, dtype=dtype).contiguous()


_CONST = {}


def _const(a, dtype, device):
    """Device copy of a small host constant (hyper-parameter arrays), cached by content so that a sweep
    issues no pageable host-to-device copy (which a CUDA-graph capture would not tolerate)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(a))
    key = ("arr", arr.shape, arr.dtype.str, hash(arr.tobytes()), str(dtype), str(device))
    t = _CONST.get(key)
    if t is None:
        if len(_CONST) > 256:
            _CONST.clear()
        t = _CONST[key] = torch.as_tensor(arr).to(device=device, dtype=dtype).contiguous()
    return t


def seed_to_u64(seed):
    """model['seed'] is a 2x uint32 key (JAX PRNGKey layout); fold it into one 64-bit Philox key."""
    if isinstance(seed, torch.Tensor):
        seed = seed.detach().cpu().numpy()
    s = np.asarray(seed).astype(np.uint64).reshape(-1)
    if s.size == 1:
        return int(s[0])
    return int((s[0] << np.uint64(32)) | (s[1] & np.uint64(0xFFFFFFFF)))


def advance_seed(seed):
    """Next sweep's key (splitmix64 step on the folded key), in the 2x uint32 layout."""
    v = (seed_to_u64(seed) + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    v ^= v >> 30
    v = (v * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    v ^= v >> 27
    v = (v * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    v ^= v >> 31
    return np.array([v >> 32, v & 0xFFFFFFFF], dtype=np.uint32)


def _mix(seed64, salt):
    return (seed64 ^ ((salt + 1) * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF


def lifted_obs_matrix(Cd, k, D):
    """Ct = (Gamma kron I_D) Cd, shape (k*D, d+1), float64 on Cd's device."""
    key = ("lift", k, D, str(Cd.device))
    lift = _CONST.get(key)
    if lift is None:         # constant of the model: built once per (k, D, device), never inside a graph capture
        Gamma = torch.as_tensor(center_embedding(k), dtype=torch.float64, device=Cd.device)
        lift = _CONST[key] = torch.kron(Gamma, torch.eye(D, dtype=torch.float64, device=Cd.device)).contiguous()
    return lift @ Cd.to(torch.float64)


def to_device_data(data, device="cuda", dtype=torch.float64):
    """`jax.device_put(data)` equivalent (fitting.py:377): Y/conf in `dtype`, mask int32."""
    out = {"Y": _dev(data["Y"], dtype, device), "mask": _dev(data["mask"], torch.int32, device)}
    if "conf" in data:
        out["conf"] = _dev(data["conf"], dtype, device)
    return out


def to_device_model(model, device="cuda", dtype=torch.float64):
    """`device_put_as_scalar(model)` equivalent (fitting.py:243).

    States and noise_prior in `dtype`, z int32, params float64, hypparams stay host scalars /
    NumPy arrays, seed stays a host uint32[2].
    """
    st, pr = model["states"], model["params"]
    return {
        "seed": np.asarray(model["seed"].cpu() if isinstance(model["seed"], torch.Tensor) else model["seed"],
                           dtype=np.uint32),
        "states": {"x": _dev(st["x"], dtype, device), "v": _dev(st["v"], dtype, device),
                   "h": _dev(st["h"], dtype, device), "s": _dev(st["s"], dtype, device),
                   "z": _dev(st["z"], torch.int32, device)},
        "params": {key: _dev(val, torch.float64, device) for key, val in pr.items()},
        "hypparams": model["hypparams"],
        # a checkpoint written without an error estimator may hold a scalar prior: the kernels index (N,T,k)
        "noise_prior": _dev(_broadcast_prior(model["noise_prior"], tuple(st["s"].shape) + (0,)), dtype, device),
    }


# ----------------------------------------------------------------------------
# discrete states
# ----------------------------------------------------------------------------
def _hmm_forward(x, mask, Ab, Q, pi, dtype):
    """Shared forward pass. Returns (filt, logZ, dims, ws)."""
    dev = x.device
    N, T, d = x.shape
    K = Ab.shape[0]
    L = Ab.shape[2] // d
    Tp = T - L
    ldT = (Tp + 7) // 8 * 8
    ldK = (K + 3) // 4 * 4
    code = _lib.dtype_code(dtype)
    esz = 4 if code == _lib.F32 else 8
    xx, AA, QQ, pp = (_dev(t, dtype, dev) for t in (x, Ab, Q, pi))
    ws = _scratch("hmm_ws", _lib.query("kpms_hmm_workspace_bytes", code, N, T, K, d, L), dev)
    W = _scratch("hmm_W", _lib.query("kpms_hmm_weights_bytes", code, N, T, K, L), dev)
    mx = _scratch("hmm_mx", N * ldT * esz, dev)
    filt = _scratch("hmm_filt", N * Tp * ldK * esz, dev)
    logZ = torch.empty(N, dtype=torch.float64, device=dev)
    sp = _lib.stream_ptr()
    _lib.call("kpms_ar_loglik", code, _lib.ptr(xx), _lib.ptr(mask), _lib.ptr(AA), _lib.ptr(QQ), N, T, d, L, K,
              ldT, _lib.ptr(W), _lib.ptr(mx), _lib.ptr(ws), sp)
    _lib.call("kpms_hmm_forward", code, _lib.ptr(W), _lib.ptr(mx), _lib.ptr(pp), N, K, Tp, ldT, _lib.ptr(filt),
              _lib.ptr(logZ), _lib.ptr(ws), d, L, sp)
    return filt, logZ, (N, K, Tp, d, L, code, pp), ws


def resample_discrete_stateseqs(x, mask, Ab, Q, pi, seed64=0, u_z=None, dtype=torch.float64, seed_dev=None, **kwargs):
    """z | x, params by HMM forward filtering / backward sampling.

    Mirrors jax_moseq.models.arhmm.resample_discrete_stateseqs.  `u_z` (N, T-L) uniforms puts
    the sampler in verification mode.  Returns (z int32 (N, T-L), logZ float64 (N,)).
    """
    filt, logZ, (N, K, Tp, d, L, code, pp), ws = _hmm_forward(x, mask, Ab, Q, pi, dtype)
    z = torch.empty((N, Tp), dtype=torch.int32, device=x.device)
    u = _dev(u_z, dtype, x.device)
    esz = 4 if code == _lib.F32 else 8
    u_scratch = None if u is not None else _scratch("hmm_u", N * Tp * esz, x.device)
    _lib.call("kpms_hmm_backward_sample", code, _lib.ptr(filt), _lib.ptr(pp), _lib.ptr(u), _lib.ptr(u_scratch),
              seed64, _lib.ptr(seed_dev), N, K, Tp, _lib.ptr(z), _lib.ptr(ws), d, L, _lib.stream_ptr())
    return z, logZ


def marginal_log_likelihood(mask, x, Ab, Q, pi, dtype=torch.float64, **kwargs):
    """Sum over chains of the HMM forward log-normaliser (argument order of fitting.py:667-673)."""
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else "cuda"
    x, Ab, Q, pi = (_dev(t, dtype, dev) for t in (x, Ab, Q, pi))
    mask = _dev(mask, torch.int32, dev)
    _, logZ, _, _ = _hmm_forward(x, mask, Ab, Q, pi, dtype)
    return logZ.sum()


def stateseq_marginals(x, mask, Ab, Q, pi, dtype=torch.float64, **kwargs):
    """Smoothed state marginals (N, T-L, K) (called as stateseq_marginals(x, mask, **params),
    fitting.py:536-538)."""
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else "cuda"
    x, Ab, Q, pi = (_dev(t, dtype, dev) for t in (x, Ab, Q, pi))
    mask = _dev(mask, torch.int32, dev)
    filt, _, (N, K, Tp, d, L, code, pp), _ = _hmm_forward(x, mask, Ab, Q, pi, dtype)
    marg = torch.empty((N, Tp, K), dtype=dtype, device=dev)
    _lib.call("kpms_hmm_smooth", code, _lib.ptr(filt), _lib.ptr(pp), N, K, Tp, _lib.ptr(marg), _lib.stream_ptr())
    return marg


# ----------------------------------------------------------------------------
# continuous states and per-frame resamplers
# ----------------------------------------------------------------------------
def _dims(Y, x):
    N, T, k, D = Y.shape
    return N, T, k, D, x.shape[-1]


def kalman_observation_records(Y, mask, v, h, s, Cd, sigmasq, d, L, K, Ct=None):
    """First stage of `resample_continuous_stateseqs` alone (C-ABI kpms_kalman_obs_info): the per-frame observation
    information, written into the sampler's workspace.  It needs the new noise scales but neither z nor the AR
    parameters, so the sweep runs it on a second stream beside the discrete-state kernels."""
    dev, dt = Y.device, Y.dtype
    N, T, k, D = Y.shape
    code = _lib.dtype_code(dt)
    if Ct is None:
        Ct = lifted_obs_matrix(Cd, k, D)
    Ctd, sg = _dev(Ct, dt, dev), _dev(sigmasq, dt, dev)
    ws = _scratch("kalman_ws", _lib.query("kpms_kalman_workspace_bytes", code, N, T, d, L, K), dev)
    _lib.call("kpms_kalman_obs_info", code, _lib.ptr(Y), _lib.ptr(mask), _lib.ptr(v), _lib.ptr(h), _lib.ptr(s),
              _lib.ptr(Ctd), _lib.ptr(sg), N, T, k, D, d, L, K, _lib.ptr(ws), _lib.stream_ptr())
    return Ctd, sg           # kept alive by the caller until the launch has been enqueued


def resample_continuous_stateseqs(Y, mask, v, h, s, z, Cd, sigmasq, Ab, Q, jitter=1e-3, seed64=0, w_x=None,
                                  Ct=None, seed_dev=None, info_ready=False, **kwargs):
    """x | rest by Kalman forward filtering / backward sampling over the lag-augmented state.

    Mirrors jax_moseq.models.keypoint_slds.resample_continuous_stateseqs; dtype follows Y.
    `w_x` (N, T-L+1, d*L) standard normals puts the sampler in verification mode.
    """
    dev, dt = Y.device, Y.dtype
    N, T, k, D = Y.shape
    d = Ab.shape[1]
    L = Ab.shape[2] // d
    code = _lib.dtype_code(dt)
    if Ct is None:
        Ct = lifted_obs_matrix(Cd, k, D)
    Ctd, sg, AA, QQ = (_dev(t, dt, dev) for t in (Ct, sigmasq, Ab, Q))
    K = Ab.shape[0]
    ws = _scratch("kalman_ws", _lib.query("kpms_kalman_workspace_bytes", code, N, T, d, L, K), dev)
    x = torch.empty((N, T, d), dtype=dt, device=dev)
    w = _dev(w_x, dt, dev)
    _lib.call("kpms_kalman_sample", code, _lib.ptr(Y), _lib.ptr(mask), _lib.ptr(v), _lib.ptr(h), _lib.ptr(s),
              _lib.ptr(z), _lib.ptr(Ctd), _lib.ptr(sg), _lib.ptr(AA), _lib.ptr(QQ), float(jitter), _lib.ptr(w),
              seed64, _lib.ptr(seed_dev), N, T, k, D, d, L, K, int(bool(info_ready)), _lib.ptr(x), _lib.ptr(ws),
              _lib.stream_ptr())
    return x


def resample_scales(Y, x, v, h, Cd, sigmasq, nu_s, s_0, seed64=0, g_s=None, Ct=None, seed_dev=None, **kwargs):
    """s | rest (scaled inverse chi-square per frame and keypoint); mirrors keypoint_slds.resample_scales."""
    dev, dt = Y.device, Y.dtype
    N, T, k, D, d = _dims(Y, x)
    code = _lib.dtype_code(dt)
    if Ct is None:
        Ct = lifted_obs_matrix(Cd, k, D)
    Ctd, sg = _dev(Ct, dt, dev), _dev(sigmasq, dt, dev)
    out = torch.empty((N, T, k), dtype=dt, device=dev)
    tape = _dev(g_s, dt, dev)          # keep converted operands alive until the launch is enqueued
    _lib.call("kpms_resample_scales", code, _lib.ptr(Y), _lib.ptr(x), _lib.ptr(v), _lib.ptr(h), _lib.ptr(Ctd),
              _lib.ptr(sg), _lib.ptr(s_0), float(nu_s), _lib.ptr(tape), seed64, _lib.ptr(seed_dev), N, T, k, D, d,
              _lib.ptr(out), _lib.stream_ptr())
    return out


def resample_heading_location(Y, mask, x, v, h, s, Cd, sigmasq, sigmasq_loc, fix_heading=False, seed64=0,
                              u_h=None, w_v=None, Ct=None, seed_dev=None, **kwargs):
    """(h, v) | rest: von Mises heading draw fused with the centroid pseudo-observation, then the
    random-walk FFBS.  Mirrors keypoint_slds.resample_heading followed by resample_location."""
    dev, dt = Y.device, Y.dtype
    N, T, k, D, d = _dims(Y, x)
    code = _lib.dtype_code(dt)
    if Ct is None:
        Ct = lifted_obs_matrix(Cd, k, D)
    Ctd, sg = _dev(Ct, dt, dev), _dev(sigmasq, dt, dev)
    ws = _scratch("headloc_ws", _lib.query("kpms_heading_location_workspace_bytes", code, N, T, D), dev)
    h_out = torch.empty((N, T), dtype=dt, device=dev)
    v_out = torch.empty((N, T, D), dtype=dt, device=dev)
    tu, tw = _dev(u_h, dt, dev), _dev(w_v, dt, dev)      # both must stay alive across the launch
    _lib.call("kpms_resample_heading_location", code, _lib.ptr(Y), _lib.ptr(mask), _lib.ptr(x), _lib.ptr(v),
              _lib.ptr(h), _lib.ptr(s), _lib.ptr(Ctd), _lib.ptr(sg), float(sigmasq_loc), int(bool(fix_heading)),
              _lib.ptr(tu), _lib.ptr(tw), seed64, _lib.ptr(seed_dev), N, T, k, D, d,
              _lib.ptr(h_out), _lib.ptr(v_out), _lib.ptr(ws), _lib.stream_ptr())
    return h_out, v_out


# ----------------------------------------------------------------------------
# sufficient statistics and parameter draws
# ----------------------------------------------------------------------------
def sufficient_statistics(x, z, mask, K, obs=None):
    """Packed float64 statistics of this rank's chains: [gram (K*F*F) | counts (K*K) | obsvar (k+1)].

    `obs` = (Y, v, h, s, Ct) adds the observation-variance sums (only needed when the global
    noise scale is resampled).  This buffer is the only thing all-reduced across GPUs.
    """
    dev = x.device
    N, T, d = x.shape
    L = T - z.shape[1]
    F = d * L + d + 1
    kk = 0 if obs is None else obs[0].shape[2] + 1
    packed = torch.zeros(K * F * F + K * K + kk, dtype=torch.float64, device=dev)
    sp = _lib.stream_ptr()
    ws = _scratch("stats_ws", _lib.query("kpms_ar_suffstats_workspace_bytes", N, T, d, L, K), dev)
    gram = packed[:K * F * F]
    _lib.call("kpms_ar_suffstats", _lib.dtype_code(x.dtype), _lib.ptr(x), _lib.ptr(z), _lib.ptr(mask), N, T, d, L,
              K, gram.data_ptr(), _lib.ptr(ws), sp)
    counts = torch.empty((K, K), dtype=torch.int32, device=dev)
    _lib.call("kpms_transition_counts", _lib.ptr(z), _lib.ptr(mask), N, T, L, K, _lib.ptr(counts), sp)
    packed[K * F * F:K * F * F + K * K] = counts.reshape(-1).to(torch.float64)
    if obs is not None:
        Y, v, h, s, Ct = obs
        k, D = Y.shape[2], Y.shape[3]
        ws2 = _scratch("obsvar_ws", _lib.query("kpms_obsvar_workspace_bytes", N, T, k), dev)
        out = packed[K * F * F + K * K:]
        Ctd = _dev(Ct, Y.dtype, dev)
        _lib.call("kpms_obsvar_suffstats", _lib.dtype_code(Y.dtype), _lib.ptr(Y), _lib.ptr(mask), _lib.ptr(x),
                  _lib.ptr(v), _lib.ptr(h), _lib.ptr(s), _lib.ptr(Ctd), N, T, k, D, d,
                  out.data_ptr(), _lib.ptr(ws2), sp)
    return packed


def unpack_statistics(packed, K, d, L):
    F = d * L + d + 1
    gram = packed[:K * F * F].reshape(K, F, F)
    counts = packed[K * F * F:K * F * F + K * K].round().to(torch.int32).reshape(K, K).contiguous()
    obsvar = packed[K * F * F + K * K:]
    return gram, counts, obsvar


def resample_ar_params(gram, nu_0, S_0, M_0, K_0, seed64=0, w_G=None, w_B=None, g_chi=None, seed_dev=None, **kwargs):
    """(Ab, Q) ~ MNIW posterior per state from the Gram matrices (mirrors arhmm.resample_ar_params)."""
    dev = gram.device
    K, F, _ = gram.shape
    d = S_0.shape[0]
    L = (F - d - 1) // d
    f64 = torch.float64
    S0, M0, K0 = (_const(t, f64, dev) for t in (S_0, M_0, K_0))
    Ab = torch.empty((K, d, d * L + 1), dtype=f64, device=dev)
    Q = torch.empty((K, d, d), dtype=f64, device=dev)
    gram = gram.contiguous()
    tG, tB, tC = _dev(w_G, f64, dev), _dev(w_B, f64, dev), _dev(g_chi, f64, dev)
    _lib.call("kpms_resample_ar_params", _lib.ptr(gram), _lib.ptr(K0), _lib.ptr(M0), _lib.ptr(S0),
              float(nu_0), _lib.ptr(tG), _lib.ptr(tB), _lib.ptr(tC), seed64, _lib.ptr(seed_dev), K, d, L, _lib.ptr(Ab),
              _lib.ptr(Q),
              _lib.stream_ptr())
    return Ab, Q


def resample_hdp_transitions(counts, betas, alpha, kappa, gamma, seed64=0, u_crp=None, u_bin=None, g_beta=None,
                             g_pi=None, seed_dev=None, **kwargs):
    """(betas, pi) for the weak-limit sticky HDP-HMM (mirrors utils.transitions.resample_hdp_transitions)."""
    dev = counts.device
    K = counts.shape[0]
    f64 = torch.float64
    ws = _scratch("trans_ws", _lib.query("kpms_transitions_workspace_bytes", K), dev)
    b_out = torch.empty(K, dtype=f64, device=dev)
    pi = torch.empty((K, K), dtype=f64, device=dev)
    b_in = _dev(betas, f64, dev)
    t1, t2, t3, t4 = (_dev(t, f64, dev) for t in (u_crp, u_bin, g_beta, g_pi))
    _lib.call("kpms_resample_hdp_transitions", _lib.ptr(counts), _lib.ptr(b_in), float(alpha),
              float(kappa), float(gamma), _lib.ptr(t1), _lib.ptr(t2), _lib.ptr(t3), _lib.ptr(t4), seed64,
              _lib.ptr(seed_dev), K,
              _lib.ptr(b_out), _lib.ptr(pi), _lib.ptr(ws), _lib.stream_ptr())
    return b_out, pi


def resample_obs_variance(obsvar, nu_sigma, sigmasq_0, D, seed64=0, g_sig=None, seed_dev=None, **kwargs):
    """sigmasq | rest from the reduced sums (mirrors keypoint_slds.resample_obs_variance)."""
    dev = obsvar.device
    k = obsvar.numel() - 1
    out = torch.empty(k, dtype=torch.float64, device=dev)
    stats, tape = obsvar.contiguous(), _dev(g_sig, torch.float64, dev)
    _lib.call("kpms_resample_obs_variance", _lib.ptr(stats), float(nu_sigma), float(sigmasq_0), int(D),
              _lib.ptr(tape), seed64, _lib.ptr(seed_dev), k, _lib.ptr(out), _lib.stream_ptr())
    return out


# ----------------------------------------------------------------------------
# the sweep
# ----------------------------------------------------------------------------
class SweepResult(dict):
    """The model dict a sweep returns (exactly the five reference keys).  `nan_flag`, when set, is a 0-d bool
    CUDA tensor "some resampled leaf holds a NaN" computed on the device as part of the sweep; util.NanGuard
    reads it instead of scanning the leaves again (the per-sweep check of fitting.py:30)."""
    nan_flag = None


def _sweep_device(Y, mask, prior, st, pr, hypparams, seed64, seed_loc, seed_dev, tp, ar_only, states_only,
                  resample_global_noise_scale, resample_local_noise_scale, fix_heading, jitter, hmm_dtype, group,
                  kD, sink=None, late=None):
    """The sweep proper on device-resident operands: sufficient statistics -> (all-reduce) -> parameter draws ->
    z -> s -> x -> (h, v).  `st` / `pr` are updated in place (dicts of tensors; the tensors themselves are never
    written).  `late(key)` returns operands that may still be in flight on the copy stream (host-operand path).
    Sequencing of jax_moseq.models.keypoint_slds.resample_model / arhmm.resample_model (fitting.py:25, :166-171)."""
    th, ah = hypparams["trans_hypparams"], hypparams["ar_hypparams"]
    oh, ch = hypparams["obs_hypparams"], hypparams["cen_hypparams"]
    K = int(th["num_states"])
    N, T = mask.shape
    d = st["x"].shape[-1]
    L = T - st["z"].shape[1]
    get = late if late is not None else (lambda key: {"Y": Y, "prior": prior}.get(key, st.get(key)))
    Ct = None if ar_only else lifted_obs_matrix(pr["Cd"], kD[0], kD[1])
    # Device-resident sweeps fork here: the noise scales and the observation records of the Kalman sampler depend on
    # neither the parameters drawn below nor z, so they run on a second stream beside the statistics / parameter /
    # HMM kernels (which leave most FP32 and memory capacity idle) and join before the filter starts.
    forked = None
    if (late is None and sink is None and not ar_only and resample_local_noise_scale and not tp
            and not (resample_global_noise_scale and not states_only) and _overlap_enabled()):
        main = torch.cuda.current_stream()
        side = _side_stream(st["x"].device, "prep")
        fork_ev = torch.cuda.Event()
        fork_ev.record(main)
        side.wait_event(fork_ev)
        with torch.cuda.stream(side):
            s_new = resample_scales(Y, st["x"], st["v"], st["h"], pr["Cd"], pr["sigmasq"], oh["nu_s"], prior,
                                    seed_loc, None, Ct=Ct, seed_dev=seed_dev)
            keep = kalman_observation_records(Y, mask, st["v"], st["h"], s_new, pr["Cd"], pr["sigmasq"], d, L, K, Ct=Ct)
            join_ev = torch.cuda.Event()
            join_ev.record(side)
        s_new.record_stream(main)
        forked = (s_new, join_ev, keep)
    if not states_only:
        obs = None
        if resample_global_noise_scale and not ar_only:
            obs = (get("Y"), get("v"), get("h"), get("s"), Ct)
        packed = sufficient_statistics(st["x"], st["z"], mask, K, obs)
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        gram, counts, obsvar = unpack_statistics(packed, K, d, L)
        # The transition draw (five short kernels on K CTAs) and the AR draw (K warps) are independent and each leaves
        # most of the device idle: the former runs on a second stream beside the latter (KPMS_PARAM_FORK=0: in line).
        # Single-process sweeps only: the fork was measured and validated without a process group (13.92 against
        # 14.02 ms at C2); sharded sweeps, whose graph also holds the all-reduce, keep the in-line order.
        trans_side = (_side_stream(st["x"].device, "trans")
                      if (_param_fork_enabled() and group is None and st["x"].is_cuda) else None)
        if trans_side is not None:
            main = torch.cuda.current_stream()
            fork_ev = torch.cuda.Event()
            fork_ev.record(main)
            trans_side.wait_event(fork_ev)
            betas_in = pr["betas"]
            with torch.cuda.stream(trans_side):
                pr["betas"], pr["pi"] = resample_hdp_transitions(
                    counts, betas_in, th["alpha"], th["kappa"], th["gamma"], seed64,
                    tp.get("u_crp"), tp.get("u_bin"), tp.get("g_beta"), tp.get("g_pi"), seed_dev=seed_dev)
                trans_done = torch.cuda.Event()
                trans_done.record(trans_side)
            for t_ in (counts, betas_in):                    # read on the side stream, owned by the main one
                if isinstance(t_, torch.Tensor) and t_.is_cuda:
                    t_.record_stream(trans_side)
            for t_ in (pr["betas"], pr["pi"]):               # and the other way round
                t_.record_stream(main)
        else:
            pr["betas"], pr["pi"] = resample_hdp_transitions(
                counts, pr["betas"], th["alpha"], th["kappa"], th["gamma"], seed64,
                tp.get("u_crp"), tp.get("u_bin"), tp.get("g_beta"), tp.get("g_pi"), seed_dev=seed_dev)
        pr["Ab"], pr["Q"] = resample_ar_params(gram, ah["nu_0"], ah["S_0"], ah["M_0"], ah["K_0"], seed64,
                                               tp.get("w_G"), tp.get("w_B"), tp.get("g_chi"), seed_dev=seed_dev)
        if trans_side is not None:
            torch.cuda.current_stream().wait_event(trans_done)
        if obs is not None:
            pr["sigmasq"] = resample_obs_variance(obsvar, oh["nu_sigma"], oh["sigmasq_0"], obs[0].shape[3], seed64,
                                                  tp.get("g_sig"), seed_dev=seed_dev)
    st["z"], _ = resample_discrete_stateseqs(st["x"], mask, pr["Ab"], pr["Q"], pr["pi"], seed_loc, tp.get("u_z"),
                                             dtype=hmm_dtype, seed_dev=seed_dev)
    if sink is not None:
        sink.emit("z", st["z"])
    if late is not None:
        for key in ("v", "h", "s"):
            val = late(key)
            if val is not None:
                st[key] = val
    if not ar_only:
        Yd, pri = get("Y"), get("prior")
        if forked is not None:
            torch.cuda.current_stream().wait_event(forked[1])
            st["s"] = forked[0]
        elif resample_local_noise_scale:
            st["s"] = resample_scales(Yd, st["x"], st["v"], st["h"], pr["Cd"], pr["sigmasq"], oh["nu_s"], pri,
                                      seed_loc, tp.get("g_s"), Ct=Ct, seed_dev=seed_dev)
        if sink is not None:
            sink.emit("s", st["s"])
        st["x"] = resample_continuous_stateseqs(Yd, mask, st["v"], st["h"], st["s"], st["z"], pr["Cd"],
                                                pr["sigmasq"], pr["Ab"], pr["Q"], jitter, seed_loc, tp.get("w_x"),
                                                Ct=Ct, seed_dev=seed_dev, info_ready=forked is not None)
        if sink is not None:
            sink.emit("x", st["x"])
        st["h"], st["v"] = resample_heading_location(Yd, mask, st["x"], st["v"], st["h"], st["s"], pr["Cd"],
                                                     pr["sigmasq"], ch["sigmasq_loc"], fix_heading, seed_loc,
                                                     tp.get("u_h"), tp.get("w_v"), Ct=Ct, seed_dev=seed_dev)


def _check_shapes(Yshape, mask, states, prior, ar_only):
    """The kernels take raw pointers: every operand's shape is checked against Y (N,T,k,D) here (a mismatched
    leaf - e.g. a scalar noise_prior from a checkpoint written without an error estimator - would otherwise be
    read out of bounds)."""
    N, T, k, D = Yshape
    want = {"x": (N, T, None), "v": (N, T, D), "h": (N, T), "s": (N, T, k)}
    if tuple(mask.shape) != (N, T):
        raise ValueError(f"mask has shape {tuple(mask.shape)}, expected {(N, T)}")
    for key, shp in want.items():
        if key not in states or states[key] is None:
            continue
        got = tuple(states[key].shape)
        if len(got) != len(shp) or any(b is not None and a != b for a, b in zip(got, shp)):
            raise ValueError(f"states['{key}'] has shape {got}, expected {tuple('d' if b is None else b for b in shp)}")
    z = states["z"]
    if z.shape[0] != N or not 0 < T - z.shape[1] < T:
        raise ValueError(f"states['z'] has shape {tuple(z.shape)}, expected (N={N}, T - nlags)")
    if not ar_only and prior is not None and tuple(prior.shape) != (N, T, k):
        raise ValueError(f"noise_prior has shape {tuple(prior.shape)}, expected {(N, T, k)} (broadcast it first)")


def _broadcast_prior(prior, Yshape):
    """noise_prior may be a scalar or (k,) in checkpoints written without an error estimator: expand to (N,T,k)."""
    N, T, k, _ = Yshape
    if prior is None:
        return None
    if not isinstance(prior, torch.Tensor):
        prior = torch.as_tensor(np.asarray(prior))
    if tuple(prior.shape) != (N, T, k):
        try:
            prior = prior.expand(N, T, k).contiguous()
        except RuntimeError:
            raise ValueError(f"noise_prior has shape {tuple(prior.shape)}, which does not broadcast to {(N, T, k)}") from None
    return prior


# ---- one CUDA graph per (data, shapes, options): the device-resident sweep replayed without host work ----------
_GRAPHS = {}
_GRAPH_DISABLED = set()
_GRAPH_LAUNCHES = [0]


def release_graphs():
    """Drops every captured sweep graph (and its static buffers).  A graph that contains the statistics all-reduce
    keeps its NCCL communicator busy: call this before `torch.distributed.destroy_process_group()`."""
    import gc
    _GRAPHS.clear()
    gc.collect()


def graph_kernel_launches():
    """Library kernels executed through graph replays so far (kpms_launch_count only sees direct launches)."""
    return _GRAPH_LAUNCHES[0]


def graphs_enabled():
    import os
    return os.environ.get("KPMS_GRAPH", "1") != "0"


def _param_fork_enabled():
    import os
    return os.environ.get("KPMS_PARAM_FORK", "1") != "0"


def _overlap_enabled():
    import os
    return os.environ.get("KPMS_OVERLAP", "0") == "1"      # measured at C2: 16.00 ms with, 15.95 ms without - off by default


def _hyp_key(hypparams):
    out = []
    for grp in ("trans_hypparams", "ar_hypparams", "obs_hypparams", "cen_hypparams"):
        for name, val in sorted(hypparams[grp].items()):
            if isinstance(val, (np.ndarray, torch.Tensor)):
                arr = np.ascontiguousarray(val.detach().cpu().numpy() if isinstance(val, torch.Tensor) else val)
                out.append((grp, name, arr.shape, hash(arr.tobytes())))
            else:
                out.append((grp, name, float(val)))
    return tuple(out)


class _SweepGraph:
    """Static operands + one captured CUDA graph of `_sweep_device`.

    The graph reads the model from static input buffers and the per-sweep Philox key from device memory, writes
    the resampled model into static output buffers, reduces the NaN flag and advances the key.  `step` copies the
    caller's model into the inputs (device to device, microseconds), replays the graph and returns fresh clones of
    the outputs, which keeps the functional contract of the reference (every sweep returns new arrays; older
    models stay valid for the pipelined NaN guard and for checkpoints)."""

    def __init__(self, dev, Y, mask, prior, states, params, hypparams, flags, group, rank, kD):
        self.dev, self.flags, self.group, self.hyp, self.kD = dev, flags, group, hypparams, kD
        self.Y, self.mask, self.prior = Y, mask, prior
        self.salt_loc = (_mix(0, rank)) & 0xFFFFFFFFFFFFFFFF
        self.st_in = {key: val.clone() for key, val in states.items()}
        self.pr_in = {key: val.clone() for key, val in params.items()}
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.seed_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.epoch = None
        self.next_seed = None                  # key the device holds after the last replay
        self.graph = None
        self.capture()

    def _run(self):
        st, pr = dict(self.st_in), dict(self.pr_in)
        _sweep_device(self.Y, self.mask, self.prior, st, pr, self.hyp, 0, self.salt_loc, self.seed_dev, {},
                      group=self.group, kD=self.kD, **self.flags)
        return st, pr

    def capture(self):
        side = _side_stream(self.dev, "capture")
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self._run()                                   # allocates every persistent scratch buffer
        side.synchronize()
        self.epoch = _SCRATCH_EPOCH[0]
        graph = torch.cuda.CUDAGraph()
        kw = {"capture_error_mode": "thread_local"} if self.group is not None else {}
        launched = _lib.launch_count()
        with torch.cuda.graph(graph, stream=side, **kw):
            st, pr = self._run()
            leaves = [t for t in list(st.values()) + list(pr.values()) if t.is_floating_point()]
            self.flag = torch.stack([torch.isnan(t).any() for t in leaves]).any()
            _lib.call("kpms_advance_seed", self.seed_dev.data_ptr(), _lib.stream_ptr())
        self.kernels = _lib.launch_count() - launched          # library kernels inside one replay
        self.st_out, self.pr_out, self.graph = st, pr, graph
        self.next_seed = None

    def step(self, seed, states, params, noise_prior):
        if self.epoch != _SCRATCH_EPOCH[0]:               # a scratch buffer was re-allocated since the capture
            self.capture()
        seed64 = seed_to_u64(seed)
        for key, val in states.items():
            self.st_in[key].copy_(val)
        for key, val in params.items():
            self.pr_in[key].copy_(val)
        if seed64 != self.next_seed:           # not the key the device advanced to: upload the caller's
            torch.cuda.current_stream(self.dev).synchronize()      # the pinned word may still be in flight
            self.seed_host[0] = seed64 - (1 << 64) if seed64 >= (1 << 63) else seed64
            self.seed_dev.copy_(self.seed_host, non_blocking=True)
        self.graph.replay()
        _GRAPH_LAUNCHES[0] += self.kernels
        st = {key: (val.clone() if val is not self.st_in[key] else states[key]) for key, val in self.st_out.items()}
        pr = {key: (val.clone() if val is not self.pr_in[key] else params[key]) for key, val in self.pr_out.items()}
        out = SweepResult(seed=advance_seed(seed), states=st, params=pr, hypparams=self.hyp, noise_prior=noise_prior)
        out.nan_flag = self.flag.clone()
        self.next_seed = seed_to_u64(out["seed"])
        return out


def _graph_for(dev, Y, mask, prior, states, params, hypparams, flags, group, rank, kD):
    key = (str(dev), Y.data_ptr() if Y is not None else 0, mask.data_ptr(), prior.data_ptr() if prior is not None else 0,
           tuple(mask.shape), tuple((k_, tuple(v.shape), str(v.dtype)) for k_, v in sorted(states.items())),
           tuple((k_, tuple(v.shape)) for k_, v in sorted(params.items())),
           tuple(sorted((k_, str(v)) for k_, v in flags.items())), _hyp_key(hypparams), id(group), rank)
    if key in _GRAPH_DISABLED:
        return None
    g = _GRAPHS.get(key)
    if g is None:
        if len(_GRAPHS) >= 4:                              # a fit alternates between at most a few option sets
            _GRAPHS.pop(next(iter(_GRAPHS)))
        try:
            g = _GRAPHS[key] = _SweepGraph(dev, Y, mask, prior, states, params, hypparams, flags, group, rank, kD)
        except Exception as e:  # noqa: BLE001 - capture is an optimisation: report once, run this key eagerly
            import warnings
            warnings.warn(f"keypoint_moseq_b200: CUDA-graph capture of the sweep failed ({e!r}); running eagerly")
            _GRAPH_DISABLED.add(key)
            return None
    return g


def resample_model(data, seed, states, params, hypparams, noise_prior, ar_only=False, states_only=False,
                   resample_global_noise_scale=False, resample_local_noise_scale=True, fix_heading=False,
                   verbose=False, jitter=1e-3, parallel_message_passing=False, draws=None,
                   hmm_dtype=torch.float64, group=None, host_out=None, graph=None, **kwargs):
    """One Gibbs sweep; same keywords and return layout as
    jax_moseq.models.keypoint_slds.resample_model.

    Extra keywords (all optional): `draws` = dict of injected tapes (verification mode, keys as
    in oracle.make_tape); `hmm_dtype` = arithmetic type of the discrete-state path (float64 keeps
    z bit-exact against a float64 reference); `group` = torch.distributed process group over which
    the chains are sharded (sufficient statistics are all-reduced once per sweep); `host_out` =
    dict of (pinned) host tensors keyed like `states`: each resampled state is copied into it on a
    side stream as soon as its sampler has finished (the caller synchronises before reading).
    Operands given as host tensors are uploaded on a copy stream in order of first use.
    `graph` = replay the sweep as ONE captured CUDA graph (default: whenever every operand is device-resident,
    no tapes are injected and KPMS_GRAPH != 0); the eager path launches the same kernels one by one and draws
    the same numbers.
    `parallel_message_passing` is accepted for signature compatibility: the backward pass is
    always parallel in time here and the filter recursion always serial.
    """
    tp = draws or {}
    dev = data["Y"].device if isinstance(data["Y"], torch.Tensor) and data["Y"].is_cuda else torch.device("cuda")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    x = states["x"]
    dt = x.dtype if isinstance(x, torch.Tensor) and x.dtype in (torch.float32, torch.float64) else torch.float64
    Yshape = tuple(data["Y"].shape)
    noise_prior_in = noise_prior
    noise_prior = _broadcast_prior(noise_prior, Yshape) if not ar_only else noise_prior
    prior_stable = noise_prior is noise_prior_in         # a broadcast copy has a new address every call
    _check_shapes(Yshape, data["mask"], states, noise_prior, ar_only)
    _lib.check_model_dims(x.shape[-1], Yshape[1] - states["z"].shape[1], params["pi"].shape[0])
    rank = 0
    if group is not None:
        import torch.distributed as dist
        rank = dist.get_rank(group)
    flags = dict(ar_only=bool(ar_only), states_only=bool(states_only),
                 resample_global_noise_scale=bool(resample_global_noise_scale),
                 resample_local_noise_scale=bool(resample_local_noise_scale), fix_heading=bool(fix_heading),
                 jitter=float(jitter), hmm_dtype=hmm_dtype)

    def resident(t, want=None):
        return isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and (want is None or t.dtype == want)

    on_device = (resident(data["mask"], torch.int32) and resident(states["z"], torch.int32)
                 and all(resident(states[key_], dt) for key_ in ("x", "v", "h", "s"))
                 and all(resident(val, torch.float64) for val in params.values())
                 and (ar_only or (resident(data["Y"], dt) and resident(noise_prior, dt))))
    use_graph = ((graphs_enabled() if graph is None else bool(graph)) and on_device and prior_stable and not tp
                 and host_out is None and not _lib.profiling())
    if use_graph:
        g = _graph_for(dev, None if ar_only else data["Y"], data["mask"], None if ar_only else noise_prior,
                       states, params, hypparams, flags, group, rank, Yshape[2:])
        if g is not None:
            out = g.step(seed, states, params, noise_prior_in)
            return out

    # ---- eager path.  Host operands are staged in order of first use on a copy stream; every sampler waits only
    # for what it reads, so the bulk of the transfer (Y, noise_prior, s) hides behind the HMM kernels.
    stage = _Stager(dev)
    stage.put("x", states["x"], dt)
    stage.put("z", states["z"], torch.int32)
    stage.put("mask", data["mask"], torch.int32)
    for key, val in params.items():
        stage.put("p:" + key, val, torch.float64)
    # the old noise scales are only read when they are not resampled (or feed the obs-variance statistics)
    s_needed = ar_only or not resample_local_noise_scale or (resample_global_noise_scale and not states_only)
    rest = [key for key in states if key not in ("x", "z") and (key != "s" or s_needed)]
    for key in rest:
        stage.put(key, states[key], dt)
    if not ar_only:
        stage.put("Y", data["Y"], dt)
        stage.put("prior", noise_prior, dt)
    mask = stage.get("mask")
    st = {"x": stage.get("x"), "z": stage.get("z")}
    pr = {key: stage.get("p:" + key) for key in params}
    sink = _HostSink(dev, host_out)
    seed64 = seed_to_u64(seed)
    seed_loc = _mix(seed64, rank)          # per-chain samplers: decorrelate shards
    late = lambda key: stage.get(key) if key in stage.items else None      # noqa: E731
    _sweep_device(None, mask, None, st, pr, hypparams, seed64, seed_loc, None, tp, group=group, kD=Yshape[2:],
                  sink=sink, late=late, **flags)
    sink.finish(st)
    return SweepResult(seed=advance_seed(seed), states=st, params=pr, hypparams=hypparams, noise_prior=noise_prior_in)
