"""Sequential stream behind the C ABI, on the GPU box: the whole-loop entry point ipc_agreement_check_stream (speculative
side-by-side solves, sequential semantics) against the one-call-per-candidate path and against the oracle / golden fixtures."""
import numpy as np
import pytest

from ipc_b200 import api, synth
from tests.golden_util import load, rel_err

pytestmark = pytest.mark.gpu
CHI2_RTOL = 1e-4


def _one_by_one(ipc, g, order):
    acc, mx, K = [], [], []
    for l in order:
        ok, ci = ipc.agreementCheck((g.loop_from[l], g.loop_to[l], g.loop_meas[l], g.loop_info[l]))
        acc.append(ok); mx.append(ci.max_chi2); K.append(ci.n_loops)
    return np.array(acc), np.array(mx), np.array(K)


@pytest.mark.parametrize("name,depth", [("stream_se2_intel.npz", 8), ("stream_se2_intel.npz", 3), ("stream_se3_sphere.npz", 8)])
def test_speculative_stream_matches_golden(gpu_lib, name, depth):
    z, g, cfg = load(name)
    o = z["order"]
    ipc = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
    ipc.set_option("stream_depth", depth)
    acc, info = ipc.agreementCheckStream(g.loop_from[o], g.loop_to[o], g.loop_meas[o], g.loop_info[o])
    assert np.array_equal(acc, z["accept"])
    assert np.array_equal(info["n_loops"], z["n_cluster"] + 1)
    assert rel_err(info["max_chi2"], z["max_chi2"]).max() < CHI2_RTOL
    assert np.array_equal(ipc.getMaxConsensusSet(), z["consensus"])
    assert np.allclose(ipc.poses(), z["poses"], atol=1e-6)
    ipc.close()


def test_speculative_stream_equals_one_by_one(gpu_lib):
    """Same verdicts, clusters, chi2 (to round-off) and final estimates whether the loop runs as n calls or as one call."""
    g, cfg = synth.make_config("intel", scale=0.6)
    o = g.time_order()
    a = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
    acc1, mx1, K1 = _one_by_one(a, g, o)
    b = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
    acc2, info = b.agreementCheckStream(g.loop_from[o], g.loop_to[o], g.loop_meas[o], g.loop_info[o])
    assert np.array_equal(acc1, acc2) and np.array_equal(K1, info["n_loops"])
    assert rel_err(info["max_chi2"], mx1).max() < 1e-9
    assert np.array_equal(a.getMaxConsensusSet(), b.getMaxConsensusSet())
    assert np.allclose(a.poses(), b.poses(), atol=1e-9)
    # an empty stream and a stream of one are fine too
    acc0, _ = b.agreementCheckStream(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros((0, 3, 3)))
    assert acc0.shape == (0,)
    a.close(); b.close()


def test_add_edge_then_check_equals_pair_batch(gpu_lib):
    """include/ipc_b200.h defines a pair check as `addEdgeToCnS(member); agreementCheck(cand)` on a fresh object
    (src/consensus.cpp:100-121, 42-75): the stateful API must give the batch verdict and chi2."""
    g, cfg = synth.make_config("intel", scale=0.3)
    mem, cnd = api.pair_checks(g)
    pairs = np.nonzero(mem >= 0)[0]
    sel = pairs[np.random.default_rng(2).choice(len(pairs), 12, replace=False)]
    batch = gpu_lib.IPC.from_graph(g, cfg)
    acc, info = batch.check_batch(mem[sel], cnd[sel])
    for k, c in enumerate(sel):
        m, cd = int(mem[c]), int(cnd[c])
        one = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
        one.addEdgeToCnS((g.loop_from[m], g.loop_to[m], g.loop_meas[m], g.loop_info[m]))
        one.addEdgeToCnS((g.loop_to[m], g.loop_from[m], g.loop_meas[m], g.loop_info[m]))      # same id pair: a no-op (:103-110)
        assert len(one.getMaxConsensusSet()) == 1
        ok, ci = one.agreementCheck((g.loop_from[cd], g.loop_to[cd], g.loop_meas[cd], g.loop_info[cd]))
        assert ok == acc[k] and ci.n_loops == info["n_loops"][k] and ci.window_len == info["window_len"][k]
        assert abs(ci.max_chi2 - info["max_chi2"][k]) <= CHI2_RTOL * max(abs(info["max_chi2"][k]), 1e-9)
        one.close()
    batch.close()
