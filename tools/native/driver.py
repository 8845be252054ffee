"""Builds and binds tools/native/caller_driver.cpp (bench / test infrastructure, a load generator): native
caller threads in front of the micro-batcher — one blocking call per query (run_callers) or the
submit / poll / take shape of the Go shim (run_async)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libcaller_driver.so")
SRC = os.path.join(HERE, "caller_driver.cpp")


def build() -> str:
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
                        SRC, "-o", LIB], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        from kektordb_b200 import ffi
        C.CDLL(ffi.LIB_PATH, mode=C.RTLD_GLOBAL)  # the driver resolves kdbgpu_* from it
        L = C.CDLL(build())
        L.kdb_run_callers.restype = C.c_int
        L.kdb_run_callers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.kdb_run_async.restype = C.c_int
        L.kdb_run_async.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _out(nq, k):
    return np.zeros((nq, k), np.uint32), np.zeros((nq, k), np.float64), np.zeros(nq, np.uint32)


def run_callers(batcher, queries: np.ndarray, k: int, ef_search: int, n_threads: int):
    """n_threads native threads answer `queries` one blocking call each.  Returns (ids, scores, counts, seconds)."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq, dim = q.shape
    ids, sc, cnt = _out(nq, k)
    secs = C.c_double(0.0)
    rc = lib().kdb_run_callers(batcher._h, q.ctypes.data_as(C.c_void_p), nq, dim, k, ef_search, n_threads,
                               ids.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                               cnt.ctypes.data_as(C.c_void_p), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"a caller saw error {rc}")
    return ids, sc, cnt, secs.value


def run_async(batcher, queries: np.ndarray, k: int, ef_search: int, n_submitters: int, window: int):
    """n_submitters native threads submit, ONE dispatcher thread polls and takes; `window` queries in flight."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq, dim = q.shape
    ids, sc, cnt = _out(nq, k)
    secs = C.c_double(0.0)
    rc = lib().kdb_run_async(batcher._h, q.ctypes.data_as(C.c_void_p), nq, dim, k, ef_search, n_submitters, window,
                             ids.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                             cnt.ctypes.data_as(C.c_void_p), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"a query saw error {rc}")
    return ids, sc, cnt, secs.value
