"""
TEST INFRASTRUCTURE -- ctypes binding of oracle/plm_oracle_c.c (the C/OpenMP
port of plmc's hot loops).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "plm_oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_max_threads.restype = ctypes.c_int32
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def hamming_counts(codes, thr, nthreads=0, rows=None):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    N, L = codes.shape
    if rows is None:
        out = np.zeros(N, dtype=np.int32)
        lib().oracle_hamming_counts(_p(codes, ctypes.c_uint8), ctypes.c_int64(N), ctypes.c_int32(L),
                                    ctypes.c_int32(thr), _p(out, ctypes.c_int32), ctypes.c_int32(nthreads))
    else:
        r0, r1 = rows
        out = np.zeros(r1 - r0, dtype=np.int32)
        lib().oracle_hamming_counts_rows(_p(codes, ctypes.c_uint8), ctypes.c_int64(N), ctypes.c_int32(L),
                                         ctypes.c_int32(thr), ctypes.c_int64(r0), ctypes.c_int64(r1),
                                         _p(out, ctypes.c_int32), ctypes.c_int32(nthreads))
    return out


def plm_eval(codes, w, x, q, lambda_h, lambda_J, precision="f32", nthreads=0):
    """Returns (fx, g, negloglk).  codes >= q are gaps under ignore_gaps."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    N, L = codes.shape
    if precision == "f32":
        dt, ct, fn = np.float32, ctypes.c_float, lib().oracle_plm_eval_f32
    else:
        dt, ct, fn = np.float64, ctypes.c_double, lib().oracle_plm_eval_f64
    w = np.ascontiguousarray(w, dtype=dt)
    x = np.ascontiguousarray(x, dtype=dt)
    n = L * q + L * (L - 1) // 2 * q * q
    assert x.size == n
    g = np.zeros(n, dtype=dt)
    fx = ctypes.c_double(0)
    nll = ctypes.c_double(0)
    rc = fn(_p(codes, ctypes.c_uint8), ctypes.c_int64(N), ctypes.c_int32(L), ctypes.c_int32(q),
            _p(w, ct), _p(x, ct), ct(lambda_h), ct(lambda_J), _p(g, ct),
            ctypes.byref(fx), ctypes.byref(nll), ctypes.c_int32(nthreads))
    if rc != 0:
        raise MemoryError("oracle_plm_eval allocation failed")
    return fx.value, g, nll.value
