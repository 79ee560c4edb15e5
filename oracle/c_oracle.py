"""ctypes binding of ``oracle/liboracle.so`` (test infrastructure; see ``oracle/__init__.py``)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "c", "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_swt2.restype = ctypes.c_int
        _LIB.oracle_maphashing.restype = ctypes.c_int
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return int(lib().oracle_num_threads())


def swt2(x, dec_lo, dec_hi, level, nthreads=0):
    """x: [B,C,H,W] uint8 or float32 -> [B,C,4,H,W] float32."""
    x = np.ascontiguousarray(x)
    assert x.ndim == 4 and x.dtype in (np.uint8, np.float32)
    b, c, h, w = x.shape
    lo = np.ascontiguousarray(dec_lo, dtype=np.float32)
    hi = np.ascontiguousarray(dec_hi, dtype=np.float32)
    out = np.empty((b, c, 4, h, w), dtype=np.float32)
    rc = lib().oracle_swt2(_p(x), ctypes.c_int(int(x.dtype == np.uint8)), _p(out), b, c, h, w, _p(lo), _p(hi),
                           int(lo.shape[0]), int(level), int(nthreads))
    if rc:
        raise ValueError(f"oracle_swt2 failed with {rc}")
    return out


def maphashing(q, ql, r, rl, topk=None, nthreads=0):
    """float32 +-1 codes; labels 2-D multi-hot (overlap) or 1-D (equality). -> (mAP, ap[Q], tsum[Q])."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    ql = np.asarray(ql)
    rl = np.asarray(rl)
    mode = 0 if (ql.ndim > 1 and rl.ndim > 1) else 1
    ql = np.ascontiguousarray(ql.reshape(q.shape[0], -1), dtype=np.float32)
    rl = np.ascontiguousarray(rl.reshape(r.shape[0], -1), dtype=np.float32)
    nq, n = q.shape[0], r.shape[0]
    ap = np.zeros(nq, dtype=np.float64)
    tsum = np.zeros(nq, dtype=np.int64)
    m = ctypes.c_double(0.0)
    rc = lib().oracle_maphashing(_p(q), _p(ql), _p(r), _p(rl), nq, n, q.shape[1], ql.shape[1], mode,
                                 ctypes.c_long(-1 if topk is None else int(topk)), _p(ap), _p(tsum),
                                 ctypes.byref(m), int(nthreads))
    if rc:
        raise ValueError(f"oracle_maphashing failed with {rc}")
    return float(m.value), ap, tsum
