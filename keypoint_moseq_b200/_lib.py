"""ctypes binding of libkpms_b200.so (C-ABI declared in include/kpms_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails,
an exception is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkpms_b200.so")

F32, F64 = 0, 1
GAMMA_TAPE = 13
VM_TAPE = 24

_vp, _i, _d, _u64, _sz = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_size_t

# name -> (restype, argtypes); mirrors include/kpms_b200.h one to one
SIGNATURES = {
    "kpms_version": (_i, []),
    "kpms_last_error": (C.c_char_p, []),
    "kpms_launch_count": (C.c_longlong, []),
    "kpms_profile_enable": (None, [_i]),
    "kpms_profile_report": (_i, [C.c_char_p, _sz]),
    "kpms_set_time_chunking": (None, [_i, _i, _d, _d]),
    "kpms_plan_chunks": (_i, [_i, _i, _i, _i]),
    "kpms_hmm_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "kpms_hmm_weights_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "kpms_ar_loglik": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kpms_hmm_forward": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "kpms_advance_seed": (_i, [_vp, _vp]),
    "kpms_hmm_backward_sample": (_i, [_i, _vp, _vp, _vp, _vp, _u64, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "kpms_hmm_smooth": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "kpms_kalman_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "kpms_kalman_sample": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _u64, _vp,
                                _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "kpms_kalman_obs_info": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "kpms_supported_dims": (_i, [_vp, _i]),
    "kpms_resample_scales": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _u64, _vp, _i, _i, _i, _i, _i,
                                  _vp, _vp]),
    "kpms_heading_location_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "kpms_resample_heading_location": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _u64, _vp,
                                            _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kpms_transition_counts": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "kpms_ar_suffstats_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "kpms_ar_suffstats": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "kpms_obsvar_workspace_bytes": (_sz, [_i, _i, _i]),
    "kpms_obsvar_suffstats": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "kpms_resample_ar_params": (_i, [_vp, _vp, _vp, _vp, _d, _vp, _vp, _vp, _u64, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "kpms_transitions_workspace_bytes": (_sz, [_i]),
    "kpms_resample_hdp_transitions": (_i, [_vp, _vp, _d, _d, _d, _vp, _vp, _vp, _vp, _u64, _vp, _i, _vp, _vp, _vp,
                                           _vp]),
    "kpms_resample_obs_variance": (_i, [_vp, _d, _d, _i, _vp, _u64, _vp, _i, _vp, _vp]),
}

_lib = None


class KpmsError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m keypoint_moseq_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the Gibbs kernels.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.float64:
        return F64
    raise TypeError(f"unsupported dtype {dt}: kernels compute in float32 or float64")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise KpmsError("kernel operands must be CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise KpmsError("kernel operands must be contiguous")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke an int-returning entry point; raise with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise KpmsError(f"{name} failed ({rc}): {lib.kpms_last_error().decode()}")


def query(name, *args):
    """Invoke a size query."""
    return int(getattr(load(), name)(*args))


def supported_dims():
    """[(latent_dim, nlags), ...] the library was compiled for (no device needed)."""
    if not _DIMS:
        lib = load()
        count = lib.kpms_supported_dims(None, 0)
        buf = (C.c_int * (2 * count))()
        lib.kpms_supported_dims(C.cast(buf, C.c_void_p), count)
        _DIMS.extend((buf[2 * i], buf[2 * i + 1]) for i in range(count))
    return list(_DIMS)


_DIMS = []


MAX_STATES = 512        # csrc/hmm.cu for num_states <= 128, csrc/hmm_wide.cuh (float64 only) up to 512


def check_model_dims(latent_dim, nlags, num_states):
    """Fail before the first sweep, with the supported set in the message, instead of inside a kernel launch."""
    pairs = supported_dims()
    if (int(latent_dim), int(nlags)) not in pairs:
        by_lag = {}
        for d, L in sorted(pairs):
            by_lag.setdefault(L, []).append(d)
        table = "; ".join(f"nlags={L}: latent_dim in {ds}" for L, ds in sorted(by_lag.items()))
        raise KpmsError(f"(latent_dim, nlags) = ({latent_dim}, {nlags}) is not compiled into libkpms_b200.so. "
                        f"Supported: {table}. Add the pair to KPMS_DL_GROUP_* in csrc/common.cuh and rebuild.")
    if not 1 <= int(num_states) <= MAX_STATES:
        raise KpmsError(f"num_states = {num_states} is outside 1..{MAX_STATES}")


def set_time_chunking(chunks=-1, warmup=-1, tol32=-1.0, tol64=-1.0):
    """Time-parallel chunking of the serial recursions (see include/kpms_b200.h)."""
    load().kpms_set_time_chunking(int(chunks), int(warmup), float(tol32), float(tol64))


def launch_count():
    """Kernels launched by the library so far (for bench.py's `gpu_launches`)."""
    return int(load().kpms_launch_count())


_PROFILING = [False]


def profile(enable):
    _PROFILING[0] = bool(enable)
    load().kpms_profile_enable(1 if enable else 0)


def profiling():
    """True while per-kernel event timing is on (event records cannot go into a graph capture)."""
    return _PROFILING[0]


def profile_report():
    """{kernel name: (total_ms, launches)} since the last report; synchronises the device."""
    buf = C.create_string_buffer(1 << 16)
    call("kpms_profile_report", buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, cnt = line.split()
        out[name] = (float(ms), int(cnt))
    return out
