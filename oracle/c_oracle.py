"""ctypes loader for oracle/ref_eval.c (ORACLE — tests / smoke / bench cpu legs only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libref_eval.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "ref_eval.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-ffp-contract=off",
                               src, "-o", _SO, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        for name in ("ref_eval_f32", "ref_eval_f64"):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                           ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.ref_eval_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(lib().ref_eval_max_threads())


def evaluate(tab, opt, prio, integer_starts=True, dtype=np.float32, nslot=8, want_plan=False, threads=0, nodes=1):
    """tab[J][S][8], opt[B][J] u8, prio[B][J] u8/u16 -> makespan[B] (+ start, mask)."""
    tab = np.ascontiguousarray(tab, dtype=dtype)
    J, S, W = tab.shape
    assert W == 8
    opt = np.ascontiguousarray(opt, dtype=np.uint8)
    assert prio.dtype in (np.uint8, np.uint16)
    prio = np.ascontiguousarray(prio)
    B = opt.shape[0]
    assert opt.shape == (B, J) and prio.shape == (B, J)
    mk = np.empty(B, dtype=dtype)
    start = np.zeros((B, J), dtype=dtype) if want_plan else None
    mask = np.zeros((B, J), dtype=np.uint32) if want_plan else None
    fn = lib().ref_eval_f32 if dtype == np.float32 else lib().ref_eval_f64
    rc = fn(tab.ctypes.data, J, S, opt.ctypes.data, prio.ctypes.data, prio.dtype.itemsize, B,
            int(bool(integer_starts)), nslot, int(nodes), mk.ctypes.data,
            start.ctypes.data if want_plan else None, mask.ctypes.data if want_plan else None,
            int(threads))
    if rc != 0:
        raise RuntimeError("ref_eval rc=%d" % rc)
    return (mk, start, mask) if want_plan else mk
