"""ctypes binding of ``libdaliid_b200.so`` (C-ABI declared in ``include/daliid_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library has not been
built (``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C
daliid_b200/csrc``) or no sm_100 GPU is visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DALIID_B200_LIB: another build of the same library (kernel-variant probes under tests/probes)
LIB_PATH = os.environ.get("DALIID_B200_LIB") or os.path.join(_HERE, "csrc", "libdaliid_b200.so")

# ---- constants mirrored from include/daliid_b200.h -------------------------------
ABI_VERSION = 1
OK, ERR_INVALID, ERR_CUDA, ERR_NO_VALID_QUERY, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4, -5
ERR_PEER_CAPACITY = -6
ERR_PEER_TIMEOUT = -7
METRICS = {"cosine": 0, "sqeuclidean": 1, "euclidean": 2, "dot": 3}
PRECISIONS = {"fp32": 0, "tf32x3": 1, "tf32": 2, "tf32c": 3, "f16x3": 4, "f16": 5}
ACCUMS = {"cy_f32": 0, "py_f64": 1}
K_NORMALIZE, K_DISTMAT, K_RANK_COUNT, K_RANK_FINALIZE, K_TOPK, K_FUSE, K_RANK_GATHER, K_RERANK = range(8)
KERNEL_SLOTS = {"normalize": 0, "distmat": 1, "rank_count": 2, "rank_finalize": 3, "topk": 4,
                "fuse": 5, "rank_gather": 6, "rerank": 7, "mrfuse": 8, "peer_exchange": 9, "h2d": 10}
CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy handle

NO_VALID_MSG = "Error: all query identities do not appear in gallery"


class DaliError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"daliid_b200 error {code}: {msg}")
        self.code = code


_lib = None
_lock = threading.Lock()

c_vp = ctypes.c_void_p
# typed names for readability of the table below; all data pointers travel as void* so that a plain
# integer address (numpy's __array_interface__, torch's data_ptr()) can be passed without building a
# ctypes pointer object per argument (that alone was ~15 us per evaluation call)
c_f32p = c_i32p = c_u32p = c_i64p = c_f64p = c_vp
i64 = ctypes.c_int64
ci = ctypes.c_int

# name -> (restype, argtypes); data pointers are passed as void* (host or device address)
_SIGNATURES = {
    "dali_abi_version": (ci, []),
    "dali_ctx_create": (ci, [ctypes.POINTER(c_vp), ci]),
    "dali_ctx_destroy": (None, [c_vp]),
    "dali_ctx_set_stream": (ci, [c_vp, c_vp]),
    "dali_ctx_get_stream": (c_vp, [c_vp]),
    "dali_last_error": (ctypes.c_char_p, [c_vp]),
    "dali_strerror": (ctypes.c_char_p, [ci]),
    "dali_ctx_h2d_streams": (ci, [c_vp]),
    "dali_ctx_timing_enable": (ci, [c_vp, ci]),
    "dali_ctx_timing_reset": (ci, [c_vp]),
    "dali_ctx_timing_read": (ci, [c_vp, ci, ctypes.POINTER(ci), c_f32p]),
    "dali_ctx_launch_count": (i64, [c_vp]),
    "dali_ctx_fallback_count": (i64, [c_vp]),
    "dali_ctx_plan_cache_enable": (ci, [c_vp, ci]),
    "dali_ctx_plan_cache_hits": (i64, [c_vp]),
    "dali_ctx_fused_count_enable": (ci, [c_vp, ci]),
    "dali_ctx_fused_count_calls": (i64, [c_vp]),
    "dali_normalize_f32": (ci, [c_vp, c_vp, i64, i64, i64, c_vp, i64, c_vp]),
    "dali_distmat_f32": (ci, [c_vp, c_vp, i64, c_vp, i64, i64, ci, ci, ci, c_vp, i64]),
    "dali_selftest_mean_division": (ci, [c_vp, ci, ctypes.POINTER(ctypes.c_uint64)]),
    "dali_distmat_fuse_mean_f32": (ci, [c_vp, c_vp, i64, c_vp, i64, i64, ci, ci, ci, c_vp, i64, c_vp, i64, ci, ci]),
    "dali_peer_create": (ci, [c_vp, ci, ci, i64, ctypes.POINTER(c_vp)]),
    "dali_peer_ipc_handle": (ci, [c_vp, c_vp]),
    "dali_peer_connect": (ci, [c_vp, c_vp]),
    "dali_peer_destroy": (None, [c_vp]),
    "dali_peer_capacity": (i64, [c_vp]),
    "dali_peer_buffer": (c_vp, [c_vp, ci]),
    "dali_peer_allreduce_i32": (ci, [c_vp, c_vp, ci, c_vp, i64]),
    "dali_peer_status": (ci, [c_vp]),
    "dali_eval_features_sharded_f32": (ci, [c_vp, c_vp, c_vp, i64, c_vp, i64, i64, i64, i64, c_i32p, c_i32p,
                                            c_i32p, c_i32p, ci, ci, ci, ci, ci, c_f32p, c_f64p, c_f64p,
                                            c_i32p, c_i64p, c_i64p]),
    "dali_rerank_f32": (ci, [c_vp, c_vp, i64, c_vp, i64, c_vp, i64, i64, i64, ci, ci, ctypes.c_double, c_vp, i64]),
    "dali_fuse_f32": (ci, [c_vp, ctypes.POINTER(c_vp), ci, ctypes.POINTER(c_vp),
                           ctypes.POINTER(c_vp), c_vp, i64, i64, i64]),
    "dali_argsort_f32": (ci, [c_vp, c_vp, i64, i64, i64, ci, c_vp]),
    "dali_roc_hist_f32": (ci, [c_vp, c_vp, i64, i64, i64, c_i32p, c_i32p, ci, ctypes.c_float, ctypes.c_float,
                               c_vp, c_vp]),
    "dali_mrfuse_f32": (ci, [c_vp, ctypes.POINTER(c_vp), ci, i64, i64, i64, ci, ci, ctypes.c_float, c_vp, i64,
                             c_vp, c_vp, c_vp]),
    "dali_eval_rank_f32": (ci, [c_vp, c_vp, i64, i64, i64, c_i32p, c_i32p, c_i32p, c_i32p, ci, ci,
                                c_f32p, c_f64p, c_f64p, c_i32p, c_i64p]),
    "dali_eval_features_f32": (ci, [c_vp, c_vp, i64, c_vp, i64, i64, c_i32p, c_i32p, c_i32p,
                                    c_i32p, ci, ci, ci, ci, ci, c_f32p, c_f64p, c_f64p, c_i32p,
                                    c_i64p, c_vp, i64]),
    "dali_topk_f32": (ci, [c_vp, c_vp, i64, i64, i64, ci, ci, c_vp, c_vp, c_vp]),
    "dali_topk_merge_f32": (ci, [c_vp, c_vp, c_vp, ci, i64, ci, ci, c_vp, c_vp]),
    "dali_topk_features_f32": (ci, [c_vp, c_vp, i64, c_vp, i64, i64, ci, ci, ci, ci, ci,
                                    ctypes.c_int32, c_vp, c_vp]),
    "dali_rank_plan_create": (ci, [c_vp, c_i32p, c_i32p, c_i32p, c_i32p, i64, i64,
                                   ctypes.POINTER(c_vp)]),
    "dali_rank_plan_destroy": (None, [c_vp]),
    "dali_rank_plan_num_matches": (i64, [c_vp]),
    "dali_rank_gather_keys": (ci, [c_vp, c_vp, c_vp, i64, i64, i64, c_vp]),
    "dali_rank_count": (ci, [c_vp, c_vp, c_vp, i64, i64, i64, c_vp, c_vp]),
    "dali_rank_finalize": (ci, [c_vp, c_vp, c_vp, c_vp, ci, ci, c_f32p, c_f64p, c_f64p, c_i32p,
                                c_i64p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load the shared library; raise loudly if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -c \"import __graft_entry__ as g; "
                    "g.build()\"` or `make -C daliid_b200/csrc`. daliid_b200 has no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so does not export it
                fn.restype = res
                fn.argtypes = args
            if lib.dali_abi_version() != ABI_VERSION:
                raise ImportError("libdaliid_b200.so ABI version mismatch; rebuild it")
            _lib = lib
    return _lib


class Context:
    """Owns one ``dali_ctx`` (one per device per host thread)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = c_vp()
        rc = self.lib.dali_ctx_create(ctypes.byref(h), int(device))
        if rc != OK:
            msg = self.lib.dali_last_error(None)
            raise DaliError(rc, (msg or b"").decode() or self.lib.dali_strerror(rc).decode())
        self.h = h
        self.device = int(device)
        self._stream = None  # raw stream last handed to the library (attach_torch_stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dali_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc == OK:
            return
        msg = (self.lib.dali_last_error(self.h) or b"").decode()
        if rc == ERR_NO_VALID_QUERY:
            raise AssertionError(NO_VALID_MSG)
        if rc == ERR_INVALID:
            raise ValueError(f"daliid_b200: {msg}")
        raise DaliError(rc, msg or self.lib.dali_strerror(rc).decode())

    def attach_torch_stream(self):
        """Issue this context's work on torch's current stream of the device."""
        import torch
        try:  # raw handle without building a torch.cuda.Stream object
            s = torch._C._cuda_getCurrentRawStream(self.device)
        except AttributeError:  # pragma: no cover
            s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            self.lib.dali_ctx_set_stream(self.h, s if s else CUDA_STREAM_LEGACY)
            self._stream = s

    # ---- timing -----------------------------------------------------------------
    def timing_enable(self, on=True):
        self.check(self.lib.dali_ctx_timing_enable(self.h, 1 if on else 0))

    def timing_reset(self):
        self.check(self.lib.dali_ctx_timing_reset(self.h))

    def timing_read(self):
        out = {}
        for name, slot in KERNEL_SLOTS.items():
            n, ms = ci(0), ctypes.c_float(0)
            self.check(self.lib.dali_ctx_timing_read(self.h, slot, ctypes.byref(n), ctypes.byref(ms)))
            out[name] = (n.value, ms.value)
        return out

    def h2d_streams(self):
        return int(self.lib.dali_ctx_h2d_streams(self.h))

    def launch_count(self):
        return int(self.lib.dali_ctx_launch_count(self.h))

    def plan_cache_enable(self, on=True):
        self.check(self.lib.dali_ctx_plan_cache_enable(self.h, 1 if on else 0))

    def plan_cache_hits(self):
        return int(self.lib.dali_ctx_plan_cache_hits(self.h))

    def fused_count_enable(self, on=True):
        self.check(self.lib.dali_ctx_fused_count_enable(self.h, 1 if on else 0))

    def fused_count_calls(self):
        return int(self.lib.dali_ctx_fused_count_calls(self.h))

    def fallback_count(self):
        return int(self.lib.dali_ctx_fallback_count(self.h))


_tls = threading.local()


def get_ctx(device=None) -> Context:
    """Per-thread, per-device cached context.  ``device=None`` -> torch's current device."""
    if device is None:
        import torch
        if not torch.cuda.is_available():
            load()  # surfaces a missing library first
            raise DaliError(ERR_CUDA, "no CUDA device visible (daliid_b200 has no CPU fallback)")
        device = torch.cuda.current_device()
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    if device not in cache:
        cache[device] = Context(device)
    return cache[device]


# ---- argument marshalling ---------------------------------------------------------
class Buf:
    """A host or device fp32/int32 buffer together with the object keeping it alive."""

    __slots__ = ("ptr", "keep", "device", "shape", "ld")

    def __init__(self, ptr, keep, device, shape, ld):
        self.ptr, self.keep, self.device, self.shape, self.ld = ptr, keep, device, shape, ld


_TORCH = None
_TORCH_DTYPES = None


def as_matrix(x, dtype=np.float32, name="array") -> Buf:
    """numpy / torch (cpu or cuda) 2-D array -> pointer + leading dimension (row-major)."""
    global _TORCH, _TORCH_DTYPES
    if _TORCH is None:
        try:
            import torch
            _TORCH, _TORCH_DTYPES = torch, {np.float32: torch.float32, np.int32: torch.int32}
        except ImportError:  # pragma: no cover
            _TORCH = False
    if _TORCH and isinstance(x, _TORCH.Tensor):
        t = x.detach() if x.requires_grad else x
        shape = t.shape
        if len(shape) != 2:
            raise ValueError(f"{name} must be 2-D")
        if t.dtype != _TORCH_DTYPES[dtype]:
            t = t.to(_TORCH_DTYPES[dtype])
        if shape[0] and shape[1]:
            s0, s1 = t.stride()
            if s1 != 1 or s0 < shape[1]:
                t = t.contiguous()
                s0 = shape[1]
        else:
            s0 = 1
        ld = max(s0 if shape[0] > 1 else shape[1], shape[1], 1)
        dev = t.device.index if t.is_cuda else None
        return Buf(t.data_ptr(), t, dev, (shape[0], shape[1]), ld)
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError(f"{name} must be 2-D")
    if a.dtype != dtype or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=dtype)
    return Buf(a.__array_interface__["data"][0], a, None, a.shape, max(a.shape[1], 1))


def as_i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def np_ptr(a):
    """Address of a C-contiguous numpy array (the caller keeps the array alive)."""
    return a.__array_interface__["data"][0]


def p_i32(a):
    return np_ptr(a)
