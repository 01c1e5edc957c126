"""ctypes binding of the plain-C decode oracle (oracle/decode_ref.c).  TEST ORACLE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_decode.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "decode_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_map_to_partition.restype = ctypes.c_int
        _lib.oracle_map_to_partition_lamb.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def map_to_partition_batch(qt, bt, dire, chroma_factor, return_leaves=False, lamb=(0.7, 0.7, 1.5, 0.3, 0.7)):
    """qt [n,8,8] f32 (ints), bt/dire [n,3,16,16] f32 -> hor,ver [n,16,16] u8, dire [n,3,16,16] i8.
    lamb: the constructor thresholds lamb1..lamb5 (Map2Partition.py:100)."""
    qt = np.ascontiguousarray(qt, np.float32).reshape(-1, 64)
    n = qt.shape[0]
    bt = np.ascontiguousarray(bt, np.float32).reshape(n, 768)
    dire = np.ascontiguousarray(dire, np.float32).reshape(n, 768)
    hor = np.zeros((n, 16, 16), np.uint8)
    ver = np.zeros((n, 16, 16), np.uint8)
    dout = np.zeros((n, 3, 16, 16), np.int8)
    leaves = np.zeros(n, np.int64)
    lam = (ctypes.c_double * 5)(*[float(x) for x in lamb])
    bad = lib().oracle_map_to_partition_lamb(
        _p(qt, ctypes.c_float), _p(bt, ctypes.c_float), _p(dire, ctypes.c_float), ctypes.c_int(n),
        ctypes.c_int(chroma_factor), lam, _p(hor, ctypes.c_uint8), _p(ver, ctypes.c_uint8),
        _p(dout, ctypes.c_int8), _p(leaves, ctypes.c_long))
    if bad:
        raise RuntimeError("oracle leaf cap exceeded on %d blocks" % bad)
    return (hor, ver, dout, leaves) if return_leaves else (hor, ver, dout)


def qt_postprocess(qt):
    """qt [n,1,8,8] f32 -> [n,1,8,8] f32 holding ints 0..3 (Metrics.eli_structual_error)."""
    q = np.ascontiguousarray(qt, np.float32).reshape(-1, 64)
    out = np.zeros_like(q)
    lib().oracle_qt_postprocess(_p(q, ctypes.c_float), ctypes.c_int(q.shape[0]), _p(out, ctypes.c_float))
    return out.reshape(-1, 1, 8, 8)
