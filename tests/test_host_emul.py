"""The CUDA check's arithmetic and Dogleg control flow (ipc_b200/csrc/chain_se2.cuh is __host__ __device__) compiled for the
CPU and compared with the golden fixtures: one thread per check (formula / control-flow regressions on a box without a
GPU) and 8 / 16 cooperating threads per check (the block-parallel decomposition; the device's warp shuffles and launch
shapes are still only exercised by the gpu-marked tests)."""
import shutil

import numpy as np
import pytest

from tests.golden_util import load, rel_err

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available to build the host emulation")


@pytest.mark.parametrize("name", ["pairs_se2_intel.npz", "pairs_se2_m3500.npz"])
@pytest.mark.parametrize("noise_eps", [0.0, 1e-13])
def test_emulated_kernel_matches_golden(name, noise_eps):
    from tests.host_emul import emul
    z, g, cfg = load(name)
    acc, info, sweeps = emul.check_batch(g, cfg, z["member"], z["cand"], noise_eps=noise_eps)
    assert np.array_equal(acc, z["accept"])
    assert rel_err(info["max_chi2"], z["max_chi2"]).max() < 1e-4
    assert rel_err(info["cand_chi2"], z["cand_chi2"]).max() < 1e-4
    if noise_eps == 0.0:     # full g2o retry semantics: same path as the oracle, round-off level agreement
        assert rel_err(info["max_chi2"], z["max_chi2"]).max() < 1e-6


def test_uniform_information_path_equals_general_path():
    from tests.host_emul import emul
    z, g, cfg = load("pairs_se2_m3500.npz")
    a1, i1, _ = emul.check_batch(g, cfg, z["member"], z["cand"], use_uni=1)
    a0, i0, _ = emul.check_batch(g, cfg, z["member"], z["cand"], use_uni=0)
    assert np.array_equal(a1, a0)
    assert rel_err(i1["max_chi2"], i0["max_chi2"]).max() < 1e-6


def test_early_accept_keeps_verdicts():
    from tests.host_emul import emul
    z, g, cfg = load("pairs_se2_m3500.npz")
    a1, i1, s1 = emul.check_batch(g, cfg, z["member"], z["cand"], early_accept=1, want_info=0)
    assert np.array_equal(a1, z["accept"])
    a0, i0, s0 = emul.check_batch(g, cfg, z["member"], z["cand"], early_accept=0, want_info=0)
    assert s1.sum() < s0.sum()


def test_fused_steepest_descent_pass_equals_separate_passes():
    """sd_fuse only changes how many passes compute (|h_gn|, b.b, b.h, bHb): the Dogleg trajectory must not move."""
    from tests.host_emul import emul
    z, g, cfg = load("pairs_se2_m3500.npz")
    out = {}
    try:
        for mode in (0, 1, 2):
            emul.lib().emul_set_sd_fuse(mode)
            out[mode] = emul.check_batch(g, cfg, z["member"], z["cand"])
    finally:
        emul.lib().emul_set_sd_fuse(2)
    for mode in (1, 2):
        assert np.array_equal(out[mode][0], out[0][0])
        assert np.array_equal(out[mode][1]["iterations"], out[0][1]["iterations"])
        assert rel_err(out[mode][1]["max_chi2"], out[0][1]["max_chi2"]).max() < 1e-9
        assert out[mode][2].sum() < out[0][2].sum()      # fewer passes over the chain


@pytest.mark.parametrize("nt", [8, 16])
@pytest.mark.parametrize("name,use_uni", [("pairs_se2_intel.npz", 0), ("pairs_se2_m3500.npz", 1), ("pairs_se2_m3500.npz", 0)])
def test_emulated_cta_decomposition_matches_golden(name, use_uni, nt):
    """The block decomposition of the kernel (thread segments, scratch slots, boundary vertices, the lagged b^T H b term,
    block collectives) on nt cooperating OS threads per check: same verdicts, chi2 at round-off distance (sums are
    re-associated, exactly as on the device)."""
    from tests.host_emul import emul
    z, g, cfg = load(name)
    assert emul.lib().emul_set_cta_threads(nt) == 0
    try:
        for fuse in (0, 1, 2):
            emul.lib().emul_set_sd_fuse(fuse)
            acc, info, _ = emul.check_batch(g, cfg, z["member"], z["cand"], use_uni=use_uni, n_threads=2 * nt)
            assert np.array_equal(acc, z["accept"])
            assert rel_err(info["max_chi2"], z["max_chi2"]).max() < 1e-4
            assert np.array_equal(info["window_len"], z["hi"] - z["lo"])
    finally:
        emul.lib().emul_set_sd_fuse(2)
        emul.lib().emul_set_cta_threads(1)


@pytest.mark.parametrize("nt", [1, 8, 16])
def test_emulated_edge_cases_match_oracle(oracle_lib, nt):
    """The edge cases of the GPU suite on the CPU: shortest windows (L = 2, most emulated threads own nothing), reversed loop
    direction, touching / identical / reversed-identical / disjoint intervals (src/consensus.cpp:157-159)."""
    import os
    from ipc_b200 import api, synth
    from tests.host_emul import emul
    g0 = synth.manhattan(200, 40, seed=9, noise_scale=0.5, reverse_frac=0.5)
    rng = np.random.default_rng(1)
    lf = [0, 5, 10, 12, 12, 20, 150, 198, 30, 60]
    lt = [2, 3, 12, 20, 20, 12, 10, 196, 60, 30]
    lm = np.array([oracle_lib.compose(2, oracle_lib.inverse(2, g0.gt[a]), g0.gt[b]) for a, b in zip(lf, lt)]) + rng.normal(size=(10, 3)) * 0.05
    g = synth.Graph(2, g0.n_poses, g0.odom_meas, g0.odom_info, np.concatenate([g0.loop_from, np.array(lf, dtype=np.int32)]),
                    np.concatenate([g0.loop_to, np.array(lt, dtype=np.int32)]), np.concatenate([g0.loop_meas, lm]),
                    np.concatenate([g0.loop_info, np.tile(g0.loop_info[0], (10, 1, 1))]), g0.n_true)
    cfg = dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)
    mem, cnd = api.pair_checks(g)
    b = g0.n_loops
    mem = np.concatenate([mem, np.array([b + 2, b + 3, b + 4, b + 0, b + 8, b + 9, b + 7], dtype=np.int32)])
    cnd = np.concatenate([cnd, np.array([b + 3, b + 4, b + 5, b + 7, b + 9, b + 8, b + 6], dtype=np.int32)])
    ptr, idx = api.checks_to_csr(mem, cnd)
    oacc, orep = oracle_lib.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count())
    assert emul.lib().emul_set_cta_threads(nt) == 0
    try:
        for use_uni in (1, 0):
            acc, info, _ = emul.check_batch(g, cfg, mem, cnd, use_uni=use_uni, n_threads=max(8, 2 * nt))
            assert np.array_equal(acc, oacc)
            assert rel_err(info["max_chi2"], orep["max_chi2"]).max() < 1e-4
            assert np.array_equal(info["n_loops"], orep["n_cluster"] + 1)
    finally:
        emul.lib().emul_set_cta_threads(1)


@pytest.mark.parametrize("nt", [1, 8, 16])
def test_global_state_tile_layout_matches_shared_state(nt):
    """State in global memory (tiles by step, StateAt<NT, true>: the one-warp-per-check and long-window kernels) against the
    shared-memory layout: the arithmetic is the same, so verdicts, iteration counts and chi2 must agree to the last bit."""
    from tests.host_emul import emul
    z, g, cfg = load("pairs_se2_m3500.npz")
    assert emul.lib().emul_set_cta_threads(nt) == 0
    try:
        for use_uni in (1, 0):
            emul.lib().emul_set_global_state(0)
            a0, i0, s0 = emul.check_batch(g, cfg, z["member"], z["cand"], use_uni=use_uni, n_threads=max(8, 2 * nt))
            emul.lib().emul_set_global_state(1)
            a1, i1, s1 = emul.check_batch(g, cfg, z["member"], z["cand"], use_uni=use_uni, n_threads=max(8, 2 * nt))
            assert np.array_equal(a1, a0) and np.array_equal(a1, z["accept"])
            assert np.array_equal(i1["iterations"], i0["iterations"]) and np.array_equal(s1, s0)
            assert np.array_equal(i1["max_chi2"], i0["max_chi2"])
    finally:
        emul.lib().emul_set_global_state(0)
        emul.lib().emul_set_cta_threads(1)
