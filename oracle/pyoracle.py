"""ctypes binding of the CPU ORACLE (oracle/libipc_oracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (ipc_b200/) never imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Report(C.Structure):
    _fields_ = [("accepted", C.c_int), ("slow_path", C.c_int), ("lo", C.c_int), ("hi", C.c_int), ("n_eset", C.c_int),
                ("n_cluster", C.c_int), ("iterations", C.c_int), ("evals", C.c_int), ("result", C.c_int), ("_pad", C.c_int),
                ("max_chi2", C.c_double), ("cand_chi2", C.c_double), ("sum_chi2", C.c_double)]


REPORT_DTYPE = np.dtype([("accepted", "i4"), ("slow_path", "i4"), ("lo", "i4"), ("hi", "i4"), ("n_eset", "i4"),
                         ("n_cluster", "i4"), ("iterations", "i4"), ("evals", "i4"), ("result", "i4"), ("_pad", "i4"),
                         ("max_chi2", "f8"), ("cand_chi2", "f8"), ("sum_chi2", "f8")])


class Config(C.Structure):
    _fields_ = [("s_factor", C.c_double), ("fast_reject_th", C.c_double), ("slow_reject_th", C.c_double),
                ("fast_reject_iter_base", C.c_int), ("slow_reject_iter_base", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libipc_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ipc_oracle_capi.cpp", "ipc_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Config)]
        _LIB.orc_destroy.argtypes = [C.c_void_p]
        _LIB.orc_set_noise_exit.argtypes = [C.c_void_p, C.c_double]
        _LIB.orc_agreement_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Report)]
        _LIB.orc_add_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _LIB.orc_remove_edge.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _LIB.orc_consensus_size.argtypes = [C.c_void_p]
        _LIB.orc_get_consensus.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_get_poses.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_final_optimize.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        _LIB.orc_final_optimize.restype = C.c_double
        _LIB.orc_check_batch.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _LIB.orc_edge_eval.argtypes = [C.c_int] + [C.c_void_p] * 6
        _LIB.orc_oplus.argtypes = [C.c_int] + [C.c_void_p] * 3
        _LIB.orc_compose.argtypes = [C.c_int] + [C.c_void_p] * 3
        _LIB.orc_inverse.argtypes = [C.c_int] + [C.c_void_p] * 2
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleIPC:
    """IPC<EDGE,VERTEX> of the reference (include/ipc/consensus.hpp:5-33) on flat arrays."""

    def __init__(self, graph, cfg: dict, noise_exit=False):
        """noise_exit: False / 0 = g2o's full retry semantics; True = shortcut with eps 1e-13; a float = that eps."""
        self.g = graph
        self.dim = graph.dim
        self.meas_w = 3 if self.dim == 2 else 7
        c = Config(cfg["s_factor"], cfg["fast_reject_th"], cfg["slow_reject_th"], cfg["fast_reject_iter_base"], cfg["slow_reject_iter_base"])
        om, oi = _f64(graph.odom_meas), _f64(graph.odom_info)
        self._h = lib().orc_create(self.dim, graph.n_poses, _p(om), _p(oi), C.byref(c))
        if noise_exit:
            lib().orc_set_noise_exit(self._h, 1e-13 if noise_exit is True else float(noise_exit))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def agreement_check(self, frm, to, meas, info):
        rep = Report()
        m, i = _f64(meas), _f64(info)
        ok = lib().orc_agreement_check(self._h, int(frm), int(to), _p(m), _p(i), C.byref(rep))
        return bool(ok), rep

    def add_edge(self, frm, to, meas, info):
        m, i = _f64(meas), _f64(info)
        lib().orc_add_edge(self._h, int(frm), int(to), _p(m), _p(i))

    def remove_edge(self, frm, to) -> bool:
        return bool(lib().orc_remove_edge(self._h, int(frm), int(to)))

    def consensus(self) -> np.ndarray:
        n = lib().orc_consensus_size(self._h)
        out = np.zeros((n, 2), dtype=np.int32)
        if n:
            lib().orc_get_consensus(self._h, _p(out))
        return out

    def poses(self) -> np.ndarray:
        out = np.zeros((self.g.n_poses, self.meas_w), dtype=np.float64)
        lib().orc_get_poses(self._h, _p(out))
        return out

    def final_optimize(self, max_iterations: int = 1000):
        """Final full-graph optimisation of simulating_incremental_data (src/simulation.cpp:50-65). Returns (chi2, iterations)."""
        it = C.c_int(0)
        chi2 = lib().orc_final_optimize(self._h, int(max_iterations), C.byref(it))
        return float(chi2), it.value

    def check_batch(self, check_ptr, check_idx, n_threads: int = 1):
        """Independent checks from the dead-reckoned state; see orc_check_batch."""
        g = self.g
        lf = np.ascontiguousarray(g.loop_from, dtype=np.int32)
        lt = np.ascontiguousarray(g.loop_to, dtype=np.int32)
        lm, li = _f64(g.loop_meas), _f64(g.loop_info)
        cp = np.ascontiguousarray(check_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(check_idx, dtype=np.int32)
        n = cp.shape[0] - 1
        acc = np.zeros(n, dtype=np.uint8)
        rep = np.zeros(n, dtype=REPORT_DTYPE)
        lib().orc_check_batch(self._h, _p(lf), _p(lt), _p(lm), _p(li), n, _p(cp), _p(ci), int(n_threads), _p(acc), _p(rep))
        return acc.astype(bool), rep

    def run_stream(self, order=None):
        """simulating_incremental_data's candidate loop (src/simulation.cpp:34-47)."""
        g = self.g
        order = g.time_order() if order is None else order
        acc = np.zeros(len(order), dtype=bool)
        reps = np.zeros(len(order), dtype=REPORT_DTYPE)
        for k, l in enumerate(order):
            ok, rep = self.agreement_check(g.loop_from[l], g.loop_to[l], g.loop_meas[l], g.loop_info[l])
            acc[k] = ok
            for name in REPORT_DTYPE.names:
                reps[k][name] = getattr(rep, name)
        return acc, reps


def edge_eval(dim, meas, xi, xj):
    d = 3 if dim == 2 else 6
    e, Ji, Jj = np.zeros(d), np.zeros((d, d)), np.zeros((d, d))
    m, a, b = _f64(meas), _f64(xi), _f64(xj)
    lib().orc_edge_eval(dim, _p(m), _p(a), _p(b), _p(e), _p(Ji), _p(Jj))
    return e, Ji, Jj


def oplus(dim, x, u):
    out = np.zeros(3 if dim == 2 else 7)
    a, b = _f64(x), _f64(u)
    lib().orc_oplus(dim, _p(a), _p(b), _p(out))
    return out


def compose(dim, a, b):
    out = np.zeros(3 if dim == 2 else 7)
    a, b = _f64(a), _f64(b)
    lib().orc_compose(dim, _p(a), _p(b), _p(out))
    return out


def inverse(dim, a):
    out = np.zeros(3 if dim == 2 else 7)
    a = _f64(a)
    lib().orc_inverse(dim, _p(a), _p(out))
    return out
