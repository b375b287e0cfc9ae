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
    "marginal_log_likelihood", "stateseq_marginals", "lifted_obs_matrix", "seed_to_u64",
    "advance_seed", "to_device_model", "to_device_data", "chunk_diagnostics",
]

_SCRATCH = {}


def _scratch(tag, nbytes, device):
    """Persistent byte workspace per (tag, device); grows monotonically."""
    key = (tag, str(device))
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
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
    Gamma = torch.as_tensor(center_embedding(k), dtype=torch.float64, device=Cd.device)
    return torch.kron(Gamma, torch.eye(D, dtype=torch.float64, device=Cd.device)) @ Cd.to(torch.float64)


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
        "noise_prior": _dev(model["noise_prior"], dtype, device),
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


def resample_discrete_stateseqs(x, mask, Ab, Q, pi, seed64=0, u_z=None, dtype=torch.float64, **kwargs):
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
              seed64, N, K, Tp, _lib.ptr(z), _lib.ptr(ws), d, L, _lib.stream_ptr())
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


def resample_continuous_stateseqs(Y, mask, v, h, s, z, Cd, sigmasq, Ab, Q, jitter=1e-3, seed64=0, w_x=None,
                                  Ct=None, **kwargs):
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
    ws = _scratch("kalman_ws", _lib.query("kpms_kalman_workspace_bytes", code, N, T, d, L), dev)
    x = torch.empty((N, T, d), dtype=dt, device=dev)
    w = _dev(w_x, dt, dev)
    _lib.call("kpms_kalman_sample", code, _lib.ptr(Y), _lib.ptr(mask), _lib.ptr(v), _lib.ptr(h), _lib.ptr(s),
              _lib.ptr(z), _lib.ptr(Ctd), _lib.ptr(sg), _lib.ptr(AA), _lib.ptr(QQ), float(jitter), _lib.ptr(w),
              seed64, N, T, k, D, d, L, _lib.ptr(x), _lib.ptr(ws), _lib.stream_ptr())
    return x


def resample_scales(Y, x, v, h, Cd, sigmasq, nu_s, s_0, seed64=0, g_s=None, Ct=None, **kwargs):
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
              _lib.ptr(sg), _lib.ptr(s_0), float(nu_s), _lib.ptr(tape), seed64, N, T, k, D, d,
              _lib.ptr(out), _lib.stream_ptr())
    return out


def resample_heading_location(Y, mask, x, v, h, s, Cd, sigmasq, sigmasq_loc, fix_heading=False, seed64=0,
                              u_h=None, w_v=None, Ct=None, **kwargs):
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
              _lib.ptr(tu), _lib.ptr(tw), seed64, N, T, k, D, d,
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


def resample_ar_params(gram, nu_0, S_0, M_0, K_0, seed64=0, w_G=None, w_B=None, g_chi=None, **kwargs):
    """(Ab, Q) ~ MNIW posterior per state from the Gram matrices (mirrors arhmm.resample_ar_params)."""
    dev = gram.device
    K, F, _ = gram.shape
    d = S_0.shape[0]
    L = (F - d - 1) // d
    f64 = torch.float64
    S0, M0, K0 = (_dev(t, f64, dev) for t in (S_0, M_0, K_0))
    Ab = torch.empty((K, d, d * L + 1), dtype=f64, device=dev)
    Q = torch.empty((K, d, d), dtype=f64, device=dev)
    gram = gram.contiguous()
    tG, tB, tC = _dev(w_G, f64, dev), _dev(w_B, f64, dev), _dev(g_chi, f64, dev)
    _lib.call("kpms_resample_ar_params", _lib.ptr(gram), _lib.ptr(K0), _lib.ptr(M0), _lib.ptr(S0),
              float(nu_0), _lib.ptr(tG), _lib.ptr(tB), _lib.ptr(tC), seed64, K, d, L, _lib.ptr(Ab), _lib.ptr(Q),
              _lib.stream_ptr())
    return Ab, Q


def resample_hdp_transitions(counts, betas, alpha, kappa, gamma, seed64=0, u_crp=None, u_bin=None, g_beta=None,
                             g_pi=None, **kwargs):
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
              float(kappa), float(gamma), _lib.ptr(t1), _lib.ptr(t2), _lib.ptr(t3), _lib.ptr(t4), seed64, K,
              _lib.ptr(b_out), _lib.ptr(pi), _lib.ptr(ws), _lib.stream_ptr())
    return b_out, pi


def resample_obs_variance(obsvar, nu_sigma, sigmasq_0, D, seed64=0, g_sig=None, **kwargs):
    """sigmasq | rest from the reduced sums (mirrors keypoint_slds.resample_obs_variance)."""
    dev = obsvar.device
    k = obsvar.numel() - 1
    out = torch.empty(k, dtype=torch.float64, device=dev)
    stats, tape = obsvar.contiguous(), _dev(g_sig, torch.float64, dev)
    _lib.call("kpms_resample_obs_variance", _lib.ptr(stats), float(nu_sigma), float(sigmasq_0), int(D),
              _lib.ptr(tape), seed64, k, _lib.ptr(out), _lib.stream_ptr())
    return out


# ----------------------------------------------------------------------------
# the sweep
# ----------------------------------------------------------------------------
def resample_model(data, seed, states, params, hypparams, noise_prior, ar_only=False, states_only=False,
                   resample_global_noise_scale=False, resample_local_noise_scale=True, fix_heading=False,
                   verbose=False, jitter=1e-3, parallel_message_passing=False, draws=None,
                   hmm_dtype=torch.float64, group=None, host_out=None, **kwargs):
    """One Gibbs sweep; same keywords and return layout as
    jax_moseq.models.keypoint_slds.resample_model.

    Extra keywords (all optional): `draws` = dict of injected tapes (verification mode, keys as
    in oracle.make_tape); `hmm_dtype` = arithmetic type of the discrete-state path (float64 keeps
    z bit-exact against a float64 reference); `group` = torch.distributed process group over which
    the chains are sharded (sufficient statistics are all-reduced once per sweep); `host_out` =
    dict of (pinned) host tensors keyed like `states`: each resampled state is copied into it on a
    side stream as soon as its sampler has finished (the caller synchronises before reading).
    Operands given as host tensors are uploaded on a copy stream in order of first use.
    `parallel_message_passing` is accepted for signature compatibility: the backward pass is
    always parallel in time here and the filter recursion always serial.
    """
    tp = draws or {}
    dev = data["Y"].device if isinstance(data["Y"], torch.Tensor) and data["Y"].is_cuda else torch.device("cuda")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    x = states["x"]
    dt = x.dtype if isinstance(x, torch.Tensor) and x.dtype in (torch.float32, torch.float64) else torch.float64
    # Host operands are staged in order of first use on a copy stream; every sampler waits only for
    # what it reads, so the bulk of the transfer (Y, noise_prior, s) hides behind the HMM kernels.
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
    th, ah = hypparams["trans_hypparams"], hypparams["ar_hypparams"]
    oh, ch = hypparams["obs_hypparams"], hypparams["cen_hypparams"]
    K = int(th["num_states"])
    N, T, k, D = data["Y"].shape
    d = st["x"].shape[-1]
    L = T - st["z"].shape[1]
    sink = _HostSink(dev, host_out)

    seed64 = seed_to_u64(seed)
    rank = 0
    if group is not None:
        import torch.distributed as dist
        rank = dist.get_rank(group)
    seed_loc = _mix(seed64, rank)          # per-chain samplers: decorrelate shards
    Ct = None if ar_only else lifted_obs_matrix(pr["Cd"], k, D)

    if not states_only:
        obs = None
        if resample_global_noise_scale and not ar_only:
            obs = (stage.get("Y"), stage.get("v"), stage.get("h"), stage.get("s"), Ct)
        packed = sufficient_statistics(st["x"], st["z"], mask, K, obs)
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        gram, counts, obsvar = unpack_statistics(packed, K, d, L)
        pr["betas"], pr["pi"] = resample_hdp_transitions(
            counts, pr["betas"], th["alpha"], th["kappa"], th["gamma"], seed64,
            tp.get("u_crp"), tp.get("u_bin"), tp.get("g_beta"), tp.get("g_pi"))
        pr["Ab"], pr["Q"] = resample_ar_params(gram, ah["nu_0"], ah["S_0"], ah["M_0"], ah["K_0"], seed64,
                                               tp.get("w_G"), tp.get("w_B"), tp.get("g_chi"))
        if obs is not None:
            pr["sigmasq"] = resample_obs_variance(obsvar, oh["nu_sigma"], oh["sigmasq_0"], D, seed64,
                                                  tp.get("g_sig"))

    st["z"], _ = resample_discrete_stateseqs(st["x"], mask, pr["Ab"], pr["Q"], pr["pi"], seed_loc, tp.get("u_z"),
                                             dtype=hmm_dtype)
    sink.emit("z", st["z"])
    for key in rest:
        st[key] = stage.get(key)
    if not ar_only:
        Y, prior = stage.get("Y"), stage.get("prior")
        if resample_local_noise_scale:
            st["s"] = resample_scales(Y, st["x"], st["v"], st["h"], pr["Cd"], pr["sigmasq"], oh["nu_s"], prior,
                                      seed_loc, tp.get("g_s"), Ct=Ct)
        sink.emit("s", st["s"])
        st["x"] = resample_continuous_stateseqs(Y, mask, st["v"], st["h"], st["s"], st["z"], pr["Cd"],
                                                pr["sigmasq"], pr["Ab"], pr["Q"], jitter, seed_loc, tp.get("w_x"),
                                                Ct=Ct)
        sink.emit("x", st["x"])
        st["h"], st["v"] = resample_heading_location(Y, mask, st["x"], st["v"], st["h"], st["s"], pr["Cd"],
                                                     pr["sigmasq"], ch["sigmasq_loc"], fix_heading, seed_loc,
                                                     tp.get("u_h"), tp.get("w_v"), Ct=Ct)
    sink.finish(st)
    return {"seed": advance_seed(seed), "states": st, "params": pr, "hypparams": hypparams,
            "noise_prior": noise_prior}
