"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C ABI
(libipc_b200.so via ipc_b200.api); the oracle / golden fixtures are only the checker.
Bar (BASELINE.json north_star): verdict bits identical, chi2 within 1e-4 relative."""
import os

import numpy as np
import pytest

from ipc_b200 import api, synth
from tests.golden_util import load, rel_err

pytestmark = pytest.mark.gpu
CHI2_RTOL = 1e-4


def _compare(acc, info, z):
    assert np.array_equal(acc, z["accept"]), f"verdict mismatches at {np.nonzero(acc != z['accept'])[0][:10]}"
    assert rel_err(info["max_chi2"], z["max_chi2"]).max() < CHI2_RTOL
    assert rel_err(info["cand_chi2"], z["cand_chi2"]).max() < CHI2_RTOL
    assert np.array_equal(info["window_len"], z["hi"] - z["lo"])


@pytest.mark.parametrize("name", ["pairs_se2_intel.npz", "pairs_se2_m3500.npz"])
@pytest.mark.parametrize("noise_exit", [1, 0])
def test_pair_batch_matches_golden(gpu_lib, name, noise_exit):
    z, g, cfg = load(name)
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    ipc.set_option("noise_exit", noise_exit)
    acc, info = ipc.check_batch(z["member"], z["cand"])
    _compare(acc, info, z)
    sl, sk, nl = ipc.last_batch_stats()
    assert sl == int((z["hi"] - z["lo"]).sum()) and nl >= 2
    ipc.close()


def test_pair_batch_matches_oracle_live(gpu_lib, oracle_lib):
    """Fresh seeded input (not a fixture) at a size the oracle finishes in seconds."""
    g, cfg = synth.make_config("m3500", scale=0.12)
    mem, cnd = api.pair_checks(g)
    sel = np.sort(np.random.default_rng(5).choice(len(cnd), 3000, replace=False))
    mem, cnd = mem[sel], cnd[sel]
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    acc, info = ipc.check_batch(mem, cnd)
    ptr, idx = api.checks_to_csr(mem, cnd)
    oacc, orep = oracle_lib.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count())
    assert np.array_equal(acc, oacc)
    assert rel_err(info["max_chi2"], orep["max_chi2"]).max() < CHI2_RTOL
    ipc.close()


def test_edge_cases(gpu_lib, oracle_lib):
    """Shortest windows (L = 2), reversed loop direction, touching intervals (no overlap -> fast path on the
    candidate alone, src/consensus.cpp:157-159), identical intervals, member listed but disjoint, empty batch."""
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
    extra_m = [b + 2, b + 3, b + 4, b + 0, b + 8, b + 9, b + 7]
    extra_c = [b + 3, b + 4, b + 5, b + 7, b + 9, b + 8, b + 6]     # touching, identical, reversed-identical, disjoint, ...
    mem = np.concatenate([mem, np.array(extra_m, dtype=np.int32)])
    cnd = np.concatenate([cnd, np.array(extra_c, dtype=np.int32)])
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    acc, info = ipc.check_batch(mem, cnd)
    ptr, idx = api.checks_to_csr(mem, cnd)
    oacc, orep = oracle_lib.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count())
    assert np.array_equal(acc, oacc)
    assert rel_err(info["max_chi2"], orep["max_chi2"]).max() < CHI2_RTOL
    assert np.array_equal(info["n_loops"], orep["n_cluster"] + 1)
    acc0, info0 = ipc.check_batch(np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert acc0.shape == (0,)
    with pytest.raises(api.IpcError):
        ipc.check_batch(np.array([-1], dtype=np.int32), np.array([g.n_loops], dtype=np.int32))
    ipc.close()


def test_full_size_properties(gpu_lib):
    """BASELINE config 2 size (M3500 + 1000 outliers): properties that need no oracle.
    (1) zero-noise true loops are all accepted with chi2 ~ 0 on every pair; (2) a pair check equals the
    fast check of the candidate when the member does not overlap; (3) results are deterministic."""
    g, cfg = synth.make_config("m3500", noise_scale=0.0)
    mem, cnd = api.pair_checks(g)
    true_pair = (cnd < g.n_true) & ((mem < 0) | (mem < g.n_true))
    sel = np.nonzero(true_pair)[0]
    sel = np.sort(np.random.default_rng(2).choice(sel, 20000, replace=False))
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    acc, info = ipc.check_batch(mem[sel], cnd[sel])
    assert acc.all() and info["max_chi2"].max() < 1e-9
    g, cfg = synth.make_config("m3500")
    ipc2 = gpu_lib.IPC.from_graph(g, cfg)
    a, b = np.minimum(g.loop_from, g.loop_to), np.maximum(g.loop_from, g.loop_to)
    rng = np.random.default_rng(3)
    c = rng.integers(0, g.n_loops, 20000).astype(np.int32)
    m = rng.integers(0, g.n_loops, 20000).astype(np.int32)
    disjoint = (np.minimum(b[m], b[c]) - np.maximum(a[m], a[c])) <= 0
    acc_p, info_p = ipc2.check_batch(m, c)
    acc_f, info_f = ipc2.check_batch(np.full_like(c, -1), c)
    assert disjoint.sum() > 1000
    assert np.array_equal(acc_p[disjoint], acc_f[disjoint])
    assert np.array_equal(info_p["max_chi2"][disjoint], info_f["max_chi2"][disjoint])
    acc_p2, info_p2 = ipc2.check_batch(m, c)
    assert np.array_equal(acc_p, acc_p2) and np.array_equal(info_p["max_chi2"], info_p2["max_chi2"])
    ipc.close(); ipc2.close()


def _unpack(rows, n):
    return ((rows[:, np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)


def test_consistency_matrix_and_greedy_consensus(gpu_lib, oracle_lib):
    """N_c x N_c matrix (diagonal = fast check, overlapping pairs = K = 2 check, others = AND of the diagonals) against the
    oracle's verdicts for every solved check, and the row-AND + popcount greedy growth against a numpy restatement."""
    g, cfg = synth.make_config("intel", scale=0.3)
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    rows, order, solved = ipc.consistency_matrix()
    n = g.n_loops
    assert np.array_equal(order, g.time_order())
    mem, cnd = api.pair_checks(g)
    assert solved == len(cnd)
    ptr, idx = api.checks_to_csr(mem, cnd)
    oacc, _ = oracle_lib.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count())
    pos = np.empty(n, dtype=np.int64); pos[order] = np.arange(n)
    want = np.zeros((n, n), dtype=bool)
    diag = np.zeros(n, dtype=bool)
    for m, c, a in zip(mem, cnd, oacc):
        if m < 0:
            diag[pos[c]] = a
    want = np.logical_and.outer(diag, diag)
    for m, c, a in zip(mem, cnd, oacc):
        if m >= 0:
            want[pos[m], pos[c]] = want[pos[c], pos[m]] = a
    want[np.arange(n), np.arange(n)] = diag
    got = _unpack(rows, n)
    assert np.array_equal(got, want)
    assert np.array_equal(got, got.T)
    in_set = ipc.greedy_consensus(rows)
    S = []
    ref = np.zeros(n, dtype=bool)
    for k in range(n):
        if want[k, k] and all(want[k, s] for s in S):
            S.append(k); ref[k] = True
    assert np.array_equal(in_set, ref)
    ipc.close()


def _run_stream(ipc, g, order):
    acc, mx, cc, K = [], [], [], []
    for l in order:
        ok, ci = ipc.agreementCheck((g.loop_from[l], g.loop_to[l], g.loop_meas[l], g.loop_info[l]))
        acc.append(ok); mx.append(ci.max_chi2); cc.append(ci.cand_chi2); K.append(ci.n_loops)
    return np.array(acc), np.array(mx), np.array(cc), np.array(K)


def test_sequential_stream_matches_golden(gpu_lib):
    """IPC::agreementCheck driven like simulating_incremental_data (src/simulation.cpp:34-47): clusters of K - 1 accepted
    loops, commit / rollback, re-dead-reckoning after an accept. Inlier set bit-identical, chi2 within 1e-4, poses equal."""
    z, g, cfg = load("stream_se2_intel.npz")
    ipc = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
    acc, mx, cc, K = _run_stream(ipc, g, z["order"])
    assert np.array_equal(acc, z["accept"])
    assert np.array_equal(K, z["n_cluster"] + 1)
    assert rel_err(mx, z["max_chi2"]).max() < CHI2_RTOL
    assert np.array_equal(ipc.getMaxConsensusSet(), z["consensus"])
    assert np.allclose(ipc.poses(), z["poses"], atol=1e-6)
    ipc.close()


def test_sequential_stream_matches_oracle_live(gpu_lib, oracle_lib):
    g, cfg = synth.make_config("intel", scale=0.5)
    oacc, orep = oracle_lib.OracleIPC(g, cfg, noise_exit=True).run_stream()
    ipc = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
    acc, mx, cc, K = _run_stream(ipc, g, g.time_order())
    assert np.array_equal(acc, oacc)
    assert rel_err(mx, orep["max_chi2"]).max() < CHI2_RTOL
    assert K.max() > 10
    # consensus-set edit API on the live object (src/consensus.cpp:77-121)
    cs = ipc.getMaxConsensusSet()
    assert ipc.removeEdgeFromCnS((cs[0][1], cs[0][0])) and not ipc.removeEdgeFromCnS((cs[0][0], cs[0][1]))
    assert len(ipc.getMaxConsensusSet()) == len(cs) - 1
    ipc.close()


@pytest.mark.parametrize("name,scale,tester,width", [("intel", 0.3, "ipc_tester_2D", 3), ("sphere", 0.05, "ipc_tester_3D", 7)])
def test_cli_end_to_end_matches_oracle_stream(gpu_lib, oracle_lib, tmp_path, name, scale, tester, width):
    """ipc_tester_2D / _3D -c cfg.yaml on a synthetic g2o file: precision / recall of the .PR file equal the oracle stream's, the
    trajectory file has one pose per vertex, and the final optimisation runs (src/simulation.cpp:50-105)."""
    import subprocess
    from ipc_b200 import g2o
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "cli")])
    g, cfg = synth.make_config(name, scale=scale)
    ds, gt, out, yml = (str(tmp_path / f) for f in ("graph.g2o", "gt.txt", "res.txt", "cfg.yaml"))
    g2o.write_g2o(g, ds); g2o.write_trajectory(g.gt, gt); g2o.write_config(yml, name, ds, gt, out, g.n_true, cfg)
    r = subprocess.run([os.path.join(root, "cli", tester), "-c", yml, "--quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    oacc, _ = oracle_lib.OracleIPC(g, cfg, noise_exit=True).run_stream()
    truth = g.time_order() < g.n_true
    tp, fp, fn = (oacc & truth).sum(), (oacc & ~truth).sum(), (~oacc & truth).sum()
    pr = open(out[:-3] + "PR").read().split()
    assert float(pr[0]) == pytest.approx(tp / max(1, tp + fp), rel=1e-5) and float(pr[1]) == pytest.approx(tp / max(1, tp + fn), rel=1e-5)
    traj = np.loadtxt(out)
    assert traj.shape == (g.n_poses, width) and np.isfinite(traj).all()
    assert f"TP {tp} FP {fp}" in r.stdout


@pytest.mark.parametrize("noise_exit", [1, 0])
def test_se3_pair_batch_matches_golden(gpu_lib, noise_exit):
    """EdgeSE3 / VertexSE3 instantiation (src/consensus.cpp:175): sphere-shaped SE(3) graph, fast + pair checks."""
    z, g, cfg = load("pairs_se3_sphere.npz")
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    ipc.set_option("noise_exit", noise_exit)
    acc, info = ipc.check_batch(z["member"], z["cand"])
    _compare(acc, info, z)
    ipc.close()


def test_se3_matches_oracle_live_and_matrix(gpu_lib, oracle_lib):
    g, cfg = synth.make_config("sphere", scale=0.08)
    mem, cnd = api.pair_checks(g)
    sel = np.sort(np.random.default_rng(7).choice(len(cnd), min(len(cnd), 1500), replace=False))
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    acc, info = ipc.check_batch(mem[sel], cnd[sel])
    ptr, idx = api.checks_to_csr(mem[sel], cnd[sel])
    oacc, orep = oracle_lib.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count())
    assert np.array_equal(acc, oacc)
    assert rel_err(info["max_chi2"], orep["max_chi2"]).max() < CHI2_RTOL
    assert acc.any() and (~acc).any()
    rows, order, solved = ipc.consistency_matrix()
    assert solved == len(cnd)
    got = _unpack(rows, g.n_loops)
    assert np.array_equal(got, got.T)
    ipc.close()


def test_launch_variants_and_global_state_mode(gpu_lib):
    """Every launch shape must give the same verdicts: (a) all checks forced through the global-state kernel (MODE 1, used for
    windows longer than shared memory holds), (b) a few alternative (threads, CTAs per SM) variants, (c) the general
    (non-uniform information) kernel on a graph that qualifies for the uniform one."""
    z, g, cfg = load("pairs_se2_m3500.npz")
    ref = None
    for opts in ({}, {"bucket0_cap": 4, "bucket1_cap": 4, "bucket2_cap": 4, "bucket3_cap": 4, "bucket4_cap": 4},
                 {"cta_per_check": 1},      # the CTA-per-check table (state in shared memory) instead of one warp per check
                 {"cta_per_check": 1, "bucket0_minb": 8, "bucket1_minb": 4, "bucket2_nt": 64, "bucket2_minb": 4},
                 {"cta_per_check": 1, "bucket0_cap": 4, "bucket1_cap": 4, "bucket2_cap": 4, "bucket3_cap": 4, "bucket4_cap": 4},   # 256 threads, global state
                 {"bucket1_cap": 96, "bucket2_nt": 64, "bucket2_minb": 4, "bucket3_nt": 64, "bucket3_minb": 4},                   # two warps per check, streamed state
                 {"use_uniform": 0},
                 {"use_uniform": 0, "bucket0_cap": 4, "bucket1_cap": 4, "bucket2_cap": 4, "bucket3_cap": 4, "bucket4_cap": 4},   # general records, streamed state
                 {"speculate": 0}):
        ipc = gpu_lib.IPC.from_graph(g, cfg)
        for k, v in opts.items():
            ipc.set_option(k, v)
        acc, info = ipc.check_batch(z["member"], z["cand"])
        _compare(acc, info, z)
        if ref is None:
            ref = info["max_chi2"].copy()
        assert rel_err(info["max_chi2"], ref).max() < 1e-6
        ipc.close()
    with pytest.raises(api.IpcError):
        ipc = gpu_lib.IPC.from_graph(g, cfg)
        ipc.set_option("bucket0_nt", 96)          # not an instantiated variant: fails loudly at launch
        ipc.check_batch(z["member"], z["cand"])


def test_sequential_stream_se3_matches_golden(gpu_lib):
    """IPC<EdgeSE3, VertexSE3>::agreementCheck stream on the sphere-shaped graph: inlier set, chi2, consensus set and final
    poses against the oracle fixture; then the final full-graph optimisation lowers chi2."""
    z, g, cfg = load("stream_se3_sphere.npz")
    ipc = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
    acc, mx, cc, K = _run_stream(ipc, g, z["order"])
    assert np.array_equal(acc, z["accept"])
    assert np.array_equal(K, z["n_cluster"] + 1)
    assert rel_err(mx, z["max_chi2"]).max() < CHI2_RTOL
    assert np.array_equal(ipc.getMaxConsensusSet(), z["consensus"])
    assert np.allclose(ipc.poses(), z["poses"], atol=1e-6)
    chi2, iters = ipc.final_optimize(1000)
    assert np.isfinite(chi2) and iters >= 1
    ipc.close()


def test_sharded_matrix_equals_device_matrix(gpu_lib):
    """The multi-GPU assembly path (ipc_b200/sharding.py; world = 1 here, world = 2 over gloo in the CPU suite) gives the same
    matrix as ipc_consistency_matrix."""
    from ipc_b200 import sharding
    g, cfg = synth.make_config("intel", scale=0.25)
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    rows, order, _ = ipc.consistency_matrix()
    M, order2 = sharding.consistency_matrix_sharded(ipc, g)
    assert np.array_equal(order, order2)
    assert np.array_equal(_unpack(rows, g.n_loops), M)
    ipc.close()
