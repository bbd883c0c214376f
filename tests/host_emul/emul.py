"""ctypes wrapper of the host emulation of the CUDA check (tests/host_emul/emul.cu) — TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

from ipc_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libemul.so")
        srcs = [os.path.join(HERE, "emul.cu")] + [os.path.join(HERE, "../../ipc_b200/csrc", f) for f in ("chain_se2.cuh", "chain_se3.cuh", "common.cuh", "host_state.hpp")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "emul.cu")])
        _LIB = C.CDLL(so)
    return _LIB


def check_batch3(g, cfg, member, cand, noise_eps=1e-13, speculate=1, early_accept=0, want_info=1, n_threads=None):
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    om, oi, lf, lt, lm, li = f64(g.odom_meas), f64(g.odom_info), i32(g.loop_from), i32(g.loop_to), f64(g.loop_meas), f64(g.loop_info)
    mb, cd = i32(member), i32(cand)
    n = len(cd)
    verdict = np.zeros(n, dtype=np.uint8)
    info = np.zeros(n, dtype=api.INFO_DTYPE)
    sweeps = np.zeros(n, dtype=np.int32)
    rc = lib().emul_check_batch3(g.n_poses, p(om), p(oi), C.c_double(cfg["s_factor"]), g.n_loops, p(lf), p(lt), p(lm), p(li), n, p(mb), p(cd),
                                 C.c_double(cfg["fast_reject_th"]), C.c_double(cfg["slow_reject_th"]), cfg["fast_reject_iter_base"],
                                 cfg["slow_reject_iter_base"], C.c_double(noise_eps), speculate, early_accept, want_info, n_threads or os.cpu_count(),
                                 p(verdict), p(info), p(sweeps))
    assert rc == 0, rc
    return verdict.astype(bool), info, sweeps


def check_batch(g, cfg, member, cand, noise_eps=1e-13, speculate=1, early_accept=0, want_info=1, use_uni=1, n_threads=None, prefix_f32=0):
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    om, oi, lf, lt, lm, li = f64(g.odom_meas), f64(g.odom_info), i32(g.loop_from), i32(g.loop_to), f64(g.loop_meas), f64(g.loop_info)
    mb, cd = i32(member), i32(cand)
    n = len(cd)
    verdict = np.zeros(n, dtype=np.uint8)
    info = np.zeros(n, dtype=api.INFO_DTYPE)
    sweeps = np.zeros(n, dtype=np.int32)
    rc = lib().emul_check_batch(g.n_poses, p(om), p(oi), C.c_double(cfg["s_factor"]), g.n_loops, p(lf), p(lt), p(lm), p(li), n, p(mb), p(cd),
                                C.c_double(cfg["fast_reject_th"]), C.c_double(cfg["slow_reject_th"]), cfg["fast_reject_iter_base"],
                                cfg["slow_reject_iter_base"], C.c_double(noise_eps), speculate, early_accept, want_info, use_uni, n_threads or os.cpu_count(),
                                p(verdict), p(info), p(sweeps), prefix_f32)
    assert rc == 0
    return verdict.astype(bool), info, sweeps
