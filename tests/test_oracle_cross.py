"""The reference pins nothing at the g2o boundary (parity unpinned), so the C++
oracle is cross-checked against the independent numpy restatement (dense solve, finite-difference
SE(3) Jacobians) and against analytic invariants."""
import numpy as np
import pytest

from ipc_b200 import synth
from oracle import np_oracle


def _stream_both(g, cfg, po, n_max=None):
    orc = po.OracleIPC(g, cfg)
    ref = np_oracle.NumpyIPC(g, cfg)
    order = g.time_order()[:n_max]
    out = []
    for l in order:
        ok, rep = orc.agreement_check(g.loop_from[l], g.loop_to[l], g.loop_meas[l], g.loop_info[l])
        ok2, mx2, cc2 = ref.agreement_check(g.loop_from[l], g.loop_to[l], g.loop_meas[l], g.loop_info[l])
        out.append((ok, ok2, rep.max_chi2, mx2, rep.cand_chi2, cc2))
    return out, orc, ref


def test_stream_se2_matches_numpy_oracle(oracle_lib):
    g = synth.add_outliers(synth.manhattan(70, 14, seed=11, noise_scale=0.5), 8, seed=12)
    cfg = dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)
    out, orc, ref = _stream_both(g, cfg, oracle_lib)
    assert any(o[0] for o in out) and any(not o[0] for o in out)
    for ok, ok2, mx, mx2, cc, cc2 in out:
        assert ok == ok2
        assert mx == pytest.approx(mx2, rel=1e-6, abs=1e-9)
        assert cc == pytest.approx(cc2, rel=1e-6, abs=1e-9)
    assert np.allclose(orc.poses(), np.array(ref.est), atol=1e-7)


def test_stream_se3_matches_numpy_oracle(oracle_lib):
    g = synth.add_outliers(synth.sphere(4, 8, seed=21, noise_scale=0.3), 6, seed=22)
    cfg = dict(s_factor=50.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=6.251, slow_reject_iter_base=100)
    out, _, _ = _stream_both(g, cfg, oracle_lib)
    assert any(o[0] for o in out) and any(not o[0] for o in out)
    for ok, ok2, mx, mx2, cc, cc2 in out:
        assert ok == ok2
        assert mx == pytest.approx(mx2, rel=1e-5, abs=1e-8)


def test_zero_noise_accepts_all_true_loops(oracle_lib):
    g = synth.manhattan(120, 30, seed=3, noise_scale=0.0)
    cfg = dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)
    acc, rep = oracle_lib.OracleIPC(g, cfg).run_stream()
    assert acc.all()
    assert rep["max_chi2"].max() < 1e-12


def test_gross_outlier_rejected_and_state_restored(oracle_lib):
    g = synth.manhattan(120, 30, seed=3, noise_scale=0.2)
    cfg = dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)
    orc = oracle_lib.OracleIPC(g, cfg)
    before = orc.poses()
    ok, rep = orc.agreement_check(3, 20, [25.0, -30.0, 2.0], g.loop_info[0])
    assert not ok and rep.max_chi2 > 11.345
    assert np.array_equal(before, orc.poses())       # restore<VERTEX>, src/consensus.cpp:63-67
    assert orc.consensus().shape[0] == 0


def test_noise_exit_shortcut_keeps_verdicts(oracle_lib):
    """The retry shortcut (DESIGN.md "Termination") must not change any verdict or chi2 beyond round-off."""
    g, cfg = synth.make_config("intel", scale=0.2)
    a0, r0 = oracle_lib.OracleIPC(g, cfg, noise_exit=False).run_stream()
    a1, r1 = oracle_lib.OracleIPC(g, cfg, noise_exit=True).run_stream()
    assert np.array_equal(a0, a1)
    assert np.allclose(r0["max_chi2"], r1["max_chi2"], rtol=1e-6, atol=1e-9)


def test_cluster_rules(oracle_lib):
    """src/consensus.cpp:157-159: overlap needs positive length; closure is transitive; identical interval intersects."""
    g = synth.manhattan(60, 6, seed=5, noise_scale=0.0)
    cfg = dict(s_factor=1.0, fast_reject_th=1e9, fast_reject_iter_base=5, slow_reject_th=1e9, slow_reject_iter_base=5)
    orc = oracle_lib.OracleIPC(g, cfg)
    I = np.eye(3)
    z = lambda a, b: oracle_lib.compose(2, oracle_lib.inverse(2, g.gt[a]), g.gt[b])
    orc.add_edge(10, 20, z(10, 20), I)
    orc.add_edge(18, 30, z(18, 30), I)
    orc.add_edge(40, 50, z(40, 50), I)
    ok, rep = orc.agreement_check(20, 25, z(20, 25), I)     # touches [10,20] only at a vertex, overlaps [18,30] -> pulls both
    assert rep.slow_path == 1 and rep.n_cluster == 2 and (rep.lo, rep.hi) == (10, 30)
    ok, rep = orc.agreement_check(30, 40, z(30, 40), I)     # touching both neighbours: no intersection -> fast path
    assert rep.slow_path == 0 and rep.n_cluster == 0 and (rep.lo, rep.hi) == (30, 40)
    ok, rep = orc.agreement_check(50, 40, z(50, 40), I)     # identical interval (reversed direction) intersects
    assert rep.slow_path == 1 and rep.n_cluster == 1
    assert orc.remove_edge(20, 10) and not orc.remove_edge(20, 10)
