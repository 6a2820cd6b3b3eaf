"""ORACLE (test infrastructure): ctypes access to the C restatement of the pinned resampling
arithmetic (oracle/pinned/mmf_pinned.c).  Build with ``make -C oracle/pinned`` (done by
``__graft_entry__.build()``)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libmmf_pinned.so")
_lib = None

MODES = {"multinomial": 1, "multinomial_fast": 2, "systematic": 3, "systematic_fast": 4}


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            import subprocess

            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_PATH)
        _lib.mmf_exp_pinned.restype = C.c_float
        _lib.mmf_exp_pinned.argtypes = [C.c_float]
        _lib.mmf_exp_pinned_array.restype = None
        _lib.mmf_exp_pinned_array.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _lib.mmf_pinned_resample.restype = C.c_int
        _lib.mmf_pinned_resample.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
    return _lib


def exp_pinned(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    load().mmf_exp_pinned_array(x.ctypes.data, y.ctypes.data, x.size)
    return y


def resample(logits: np.ndarray, uniforms: np.ndarray, mode: str, num_samples: int = None, return_cdf: bool = False):
    """logits (N, M) fp32; uniforms float64 (N, S) [multinomial*] or (N,) [systematic*] -> int64 (N, S)."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    N, M = logits.shape
    code = MODES[mode]
    if code >= 3:
        assert uniforms.shape == (N,)
        S = M if num_samples is None else num_samples
    else:
        assert uniforms.ndim == 2 and uniforms.shape[0] == N
        S = uniforms.shape[1]
    idx = np.empty((N, S), dtype=np.int64)
    cdf = np.empty((N, M), dtype=np.float32) if return_cdf else None
    rc = load().mmf_pinned_resample(N, M, S, logits.ctypes.data, code, uniforms.ctypes.data, idx.ctypes.data,
                                    cdf.ctypes.data if return_cdf else None)
    assert rc == 0
    return (idx, cdf) if return_cdf else idx
