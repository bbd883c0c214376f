"""Generates the golden fixtures in this directory from the CPU oracle (oracle/ipc_oracle.hpp).

The reference ships no golden vectors and cannot be built offline (parity unpinned),
so these fixtures freeze the ORACLE's outputs on seeded synthetic graphs: they guard the oracle against
regressions and give the GPU tests inputs + expected outputs that need no oracle build.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ipc_b200 import api, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def graph_arrays(g):
    return dict(dim=g.dim, n_poses=g.n_poses, odom_meas=g.odom_meas, odom_info=g.odom_info, loop_from=g.loop_from, loop_to=g.loop_to,
                loop_meas=g.loop_meas, loop_info=g.loop_info, n_true=g.n_true)


def cfg_array(cfg):
    return np.array([cfg["s_factor"], cfg["fast_reject_th"], cfg["slow_reject_th"], cfg["fast_reject_iter_base"], cfg["slow_reject_iter_base"]])


def make_pairs(name, scale, n_max, out, seed=0):
    g, cfg = synth.make_config(name, scale=scale)
    mem, cnd = api.pair_checks(g)
    if len(cnd) > n_max:
        sel = np.sort(np.random.default_rng(seed).choice(len(cnd), n_max, replace=False))
        mem, cnd = mem[sel], cnd[sel]
    ptr, idx = api.checks_to_csr(mem, cnd)
    acc, rep = po.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count())
    np.savez_compressed(os.path.join(HERE, out), cfg=cfg_array(cfg), member=mem, cand=cnd, accept=acc, max_chi2=rep["max_chi2"],
                        cand_chi2=rep["cand_chi2"], sum_chi2=rep["sum_chi2"], lo=rep["lo"], hi=rep["hi"], **graph_arrays(g))
    print(out, len(cnd), "checks", int(acc.sum()), "accepted")


def make_stream(name, scale, out):
    g, cfg = synth.make_config(name, scale=scale)
    orc = po.OracleIPC(g, cfg)
    acc, rep = orc.run_stream()
    np.savez_compressed(os.path.join(HERE, out), cfg=cfg_array(cfg), order=g.time_order(), accept=acc, max_chi2=rep["max_chi2"],
                        cand_chi2=rep["cand_chi2"], slow=rep["slow_path"], n_cluster=rep["n_cluster"], lo=rep["lo"], hi=rep["hi"],
                        consensus=orc.consensus(), poses=orc.poses(), **graph_arrays(g))
    print(out, len(acc), "candidates", int(acc.sum()), "accepted")


if __name__ == "__main__":
    make_pairs("intel", 0.15, 1500, "pairs_se2_intel.npz")
    make_pairs("m3500", 0.06, 1500, "pairs_se2_m3500.npz")
    make_stream("intel", 0.15, "stream_se2_intel.npz")
    make_pairs("sphere", 0.05, 800, "pairs_se3_sphere.npz")
    make_stream("sphere", 0.05, "stream_se3_sphere.npz")
