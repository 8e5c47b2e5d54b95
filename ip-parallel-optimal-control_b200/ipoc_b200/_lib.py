"""ctypes binding of the C ABI in include/ipoc.h (libipoc.so, built in-tree by
`__graft_entry__.build()`).  There is NO CPU fallback: if the shared library is missing or a
tensor is not on a CUDA device, the call fails loudly."""
import ctypes
import os
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libipoc.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_SZ = ctypes.c_size_t

_SIGS = {
    "ipoc_strerror": (ctypes.c_char_p, [_I]),
    "ipoc_version": (_I, []),
    "ipoc_supported": (_I, [_I, _I]),
    "ipoc_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "ipoc_set_tuning": (None, [_I, _I, _I]),
    "ipoc_set_literal_lqt": (None, [_I]),
    "ipoc_set_hier": (None, [_I, _I, _I]),
    "ipoc_workspace_init": (_I, [_P, _SZ, _P]),
    "ipoc_set_affine_occupancy": (None, [_I]),
    "ipoc_costates_f64": (_I, [_I] * 4 + [_P] * 7 + [_P, _SZ, _P]),
    "ipoc_newton_attempt_f64": (_I, [_I] * 5 + [_P] * 29 + [_P, _SZ, _P]),
    "ipoc_launch_count": (ctypes.c_ulonglong, []),
    "ipoc_carry_doubles": (_I, [_I, _I]),
    "ipoc_newton_step_f64": (_I, [_I] * 4 + [_P] * 13 + [_P, _SZ, _P]),
    "ipoc_lqt_bwd_f64": (_I, [_I] * 4 + [_P] * 16 + [_P, _SZ, _P]),
    "ipoc_lqt_fwd_f64": (_I, [_I] * 4 + [_P] * 8 + [_P, _SZ, _P]),
    "ipoc_affine_scan_f64": (_I, [_I] * 5 + [_P] * 4 + [_P, _SZ, _P]),
    "ipoc_reductions_f64": (_I, [_I] * 4 + [_P] * 8 + [_P, _SZ, _P]),
    "ipoc_accept_update_f64": (_I, [_I] + [_P] * 10 + [_P]),
    "ipoc_attempt_begin_f64": (_I, [_I] + [_P] * 5 + [_P]),
    "ipoc_trial_point_f64": (_I, [_I] * 4 + [_P] * 6 + [_P]),
    "ipoc_attempt_commit_f64": (_I, [_I] * 4 + [_P] * 8 + [_I, _P]),
    "ipoc_newton_advance_f64": (_I, [_I] * 4 + [_P] * 10 + [ctypes.c_double, _I, _P]),
    "ipoc_attempt_finish_f64": (_I, [_I] + [_P] * 15 + [ctypes.c_double, _I, _I, _P]),
    "ipoc_masked_copy_f64": (_I, [_I] * 4 + [_P] * 5 + [_P]),
    "ipoc_plant_attempt_finish_f64": (_I, [_I, _I, _I, ctypes.c_double, ctypes.c_double] + [_P] * 18
                                      + [ctypes.c_double, _I, _I, _P, _P, _P, ctypes.c_size_t, _P]),
    "ipoc_lqr_params_f64": (_I, [_I] * 4 + [_P] * 13 + [_P]),
    "ipoc_newton_bwd_reduce_f64": (_I, [_I] * 3 + [_P] * 8 + [_P, _SZ, _P]),
    "ipoc_newton_bwd_apply_f64": (_I, [_I] * 5 + [_P] * 14 + [_P, _SZ, _P]),
    "ipoc_newton_fwd_apply_f64": (_I, [_I] * 5 + [_P] * 7 + [_P, _SZ, _P]),
    "ipoc_affine_reduce_f64": (_I, [_I] * 4 + [_P] * 3 + [_P, _SZ, _P]),
    "ipoc_affine_apply_f64": (_I, [_I] * 6 + [_P] * 5 + [_P, _SZ, _P]),
    "ipoc_newton_step_host_scratch_bytes": (_SZ, [_I] * 4),
    "ipoc_newton_step_host_f64": (_I, [_I] * 4 + [_P] * 11 + [_P, _SZ, _P]),
    "ipoc_plant_dims": (_I, [_I] + [ctypes.POINTER(ctypes.c_int)] * 3),
    "ipoc_plant_derivatives_f64": (_I, [_I, _I, _I, ctypes.c_double, ctypes.c_double] + [_P] * 14 + [_P]),
    "ipoc_plant_linearize_f64": (_I, [_I, _I, _I, ctypes.c_double, ctypes.c_double] + [_P] * 9 + [_P]),
    "ipoc_plant_take_linearize_f64": (_I, [_I, _I, _I, ctypes.c_double, ctypes.c_double] + [_P] * 11 + [_P]),
    "ipoc_plant_hamiltonian_f64": (_I, [_I, _I, _I, ctypes.c_double, ctypes.c_double] + [_P] * 9 + [_P]),
    "ipoc_plant_cost_f64": (_I, [_I, _I, _I, ctypes.c_double, ctypes.c_double] + [_P] * 6 + [_P, ctypes.c_size_t, _P]),
    "ipoc_plant_cost_workspace_bytes": (ctypes.c_size_t, [_I, _I]),
    "ipoc_plant_rollout_f64": (_I, [_I, _I, _I, ctypes.c_double] + [_P] * 3 + [_P]),
    "ipoc_plant_rollout_lin_f64": (_I, [_I, _I, _I, ctypes.c_double] + [_P] * 6 + [_P]),
    "ipoc_profile_begin": (_I, [_P]),
    "ipoc_profile_end": (_I, [ctypes.c_char_p, _SZ, ctypes.POINTER(ctypes.c_float), _I]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

WS_NEWTON_STEP, WS_LQT_BWD, WS_LQT_FWD, WS_AFFINE_SCAN, WS_REDUCTIONS, WS_COSTATES, WS_NEWTON_ATTEMPT = range(7)
CARRY_RICCATI, CARRY_AFFINE = 0, 1

_lib = None


class IpocError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IpocError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise IpocError(f"ipoc error {rc}: {lib().ipoc_strerror(rc).decode()}")


def ptr(t):
    """Device pointer of a contiguous float64/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise IpocError("ipoc kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise IpocError("ipoc kernels need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def dev_f64(t, device=None):
    """float64, contiguous, 16-byte aligned CUDA view/copy of t."""
    t = torch.as_tensor(t)
    if device is not None and t.device != torch.device(device):
        t = t.to(device)
    if not t.is_cuda:
        raise IpocError("ipoc kernels need CUDA tensors; there is no CPU fallback")
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_ws_cache = {}
_ws_retired = []   # never freed; a buffer is only retired when it at least doubles, so the total stays < 2x the live size


def workspace(kind, N, nx, nu, batch, device):
    """Scratch tensor for one C-ABI call, grown on demand and reused (the C library owns nothing).

    One buffer per (device, kind, CURRENT STREAM): calls enqueued on different streams — including a CUDA graph
    captured on one stream and eager calls on another — never share scratch, which makes this binding as
    re-entrant as the C ABI itself (a workspace must not be used by two calls at the same time).  Buffers are
    zero-filled at allocation: the control block of a scan workspace must start zeroed (include/ipoc.h).
    Graph captures in this package warm up and capture on the SAME side stream (`torch.cuda.graph(g, stream=side)`), so
    the buffer exists before the capture starts — allocated inside a capture, its zero-fill would become a node of
    the graph and run on every replay."""
    need = lib().ipoc_workspace_bytes(kind, N, nx, nu, batch)
    if need == 0:
        raise IpocError(f"unsupported problem size (nx={nx}, nu={nu}): no kernel instantiated, no CPU fallback")
    dev = torch.device(device)
    key = (dev.index or 0, kind, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < need:
        if buf is not None:
            # captured CUDA graphs (graphed.py, batched.py, sharded.py) hold the old buffer's address:
            # retire it instead of freeing it, so that replaying them can never touch someone else's memory
            _ws_retired.append(buf)
        buf = torch.zeros(int(need * 2) + 1024, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf, need


def require_supported(nx, nu):
    if not lib().ipoc_supported(nx, nu):
        raise IpocError(f"unsupported (nx={nx}, nu={nu}): no kernel instantiated and there is no CPU fallback")


_tuning = (0, 0, 0)
_knob_lock = __import__("threading").RLock()


class tuning:
    """`with tuning(leaf_chunk=T): ...` — set the scan-plan knobs for the duration of a block and restore them, under a
    process-wide lock (the knobs are globals of the C library: a concurrent caller must not see a half-changed plan)."""

    def __init__(self, leaf_chunk=0, mid_fanin=0, top_max=0):
        self.new = (leaf_chunk, mid_fanin, top_max)

    def __enter__(self):
        _knob_lock.acquire()
        self.prev = set_tuning(*self.new)
        return self

    def __exit__(self, *exc):
        set_tuning(*self.prev)
        _knob_lock.release()
        return False


def set_tuning(leaf_chunk=0, mid_fanin=0, top_max=0):
    """Scan-plan knobs of the library (0 = default); remembered so callers can restore them.  The knobs
    (this, ipoc_set_hier, ipoc_set_literal_lqt) are PROCESS-GLOBAL test / experiment switches: do not flip them
    while another thread is inside the library (a plan must be carved with the knobs it was sized with)."""
    global _tuning
    prev = _tuning
    lib().ipoc_set_tuning(int(leaf_chunk), int(mid_fanin), int(top_max))
    _tuning = (int(leaf_chunk), int(mid_fanin), int(top_max))
    return prev
