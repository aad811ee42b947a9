"""ctypes wrapper of oracle/ref_port.c (OpenMP restatement of the reference's alt_proj path).

TEST / BASELINE INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libref_port.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            import subprocess
            subprocess.check_call(["bash", os.path.join(_HERE, "build_port.sh")])
        _lib = C.CDLL(LIB)
        _lib.lfe_port_fit.restype = C.c_int
        _lib.lfe_port_threads.restype = C.c_int
    return _lib


def threads() -> int:
    return int(_load().lfe_port_threads())


def set_threads(t: int) -> int:
    """Use ``t`` OpenMP threads from now on (overrides an inherited OMP_NUM_THREADS); -> the count in effect."""
    _load().lfe_port_set_threads(C.c_int(int(t)))
    return threads()


def fit(y, X: list, fe: list, cluster=None, tol=1e-6, max_iter=50, vcov="iid", ssc=True):
    """alt_proj fit on all host cores.  X: list of float64 arrays, fe: list of int64 id arrays."""
    lib = _load()
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in [y] + list(X)]
    fes = [np.ascontiguousarray(f, dtype=np.int64) for f in fe]
    k, F, n = len(X), len(fes), cols[0].shape[0]
    cp = (C.c_void_p * (k + 1))(*[c.ctypes.data for c in cols])
    fp = (C.c_void_p * F)(*[f.ctypes.data for f in fes])
    vk = {"iid": 0, "hc1": 1, "cluster": 2}[vcov.lower()]
    cl = np.ascontiguousarray(cluster, dtype=np.int64) if vk == 2 else None
    beta, se, info = np.zeros(k), np.zeros(k), np.zeros(6)
    rc = lib.lfe_port_fit(cp, C.c_int(k), fp, C.c_int(F), C.c_void_p(cl.ctypes.data if cl is not None else None),
                          C.c_int64(n), C.c_double(tol), C.c_int(max_iter), C.c_int(vk), C.c_int(1 if ssc else 0),
                          C.c_void_p(beta.ctypes.data), C.c_void_p(se.ctypes.data), C.c_void_p(info.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"lfe_port_fit failed with {rc}")
    return dict(beta=beta, se=se, iterations=int(info[0]), n_obs=int(info[1]), df_resid=int(info[2]),
                n_clusters=int(info[3]), rss=float(info[4]), max_mean=float(info[5]))
