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
