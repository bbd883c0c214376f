"""Host-side logic: candidate ordering (src/simulation.cpp:26), pair enumeration (src/consensus.cpp:157-159),
synthetic generators and the outlier distribution of scripts/generateDataset.py:188-246."""
import numpy as np

from ipc_b200 import api, synth


def test_time_order_is_stable_by_max_id():
    g, _ = synth.make_config("intel", scale=0.2)
    o = g.time_order()
    mx = np.maximum(g.loop_from, g.loop_to)[o]
    assert (np.diff(mx) >= 0).all()
    same = np.nonzero(np.diff(mx) == 0)[0]
    assert (o[same] < o[same + 1]).all()


def test_pair_checks_matches_bruteforce():
    g, _ = synth.make_config("intel", scale=0.1)
    mem, cnd = api.pair_checks(g)
    o = g.time_order()
    a, b = np.minimum(g.loop_from, g.loop_to), np.maximum(g.loop_from, g.loop_to)
    want = {(-1, int(c)) for c in o}
    for j in range(len(o)):
        for i in range(j):
            if min(b[o[i]], b[o[j]]) - max(a[o[i]], a[o[j]]) > 0:
                want.add((int(o[i]), int(o[j])))
    assert set(zip(mem.tolist(), cnd.tolist())) == want
    ptr, idx = api.checks_to_csr(mem, cnd)
    assert ptr[-1] == len(idx) == len(cnd) + (mem >= 0).sum()
    assert (idx[ptr[1:] - 1] == cnd).all()


def test_outliers_follow_reference_script_rules():
    g, _ = synth.make_config("m3500", scale=0.1)
    lf, lt = g.loop_from[g.n_true:], g.loop_to[g.n_true:]
    assert (lf < lt).all() and (lt - lf >= 2).all() and lt.max() <= g.n_poses - 1
    assert np.allclose(g.loop_info[g.n_true:], g.loop_info[0])
    assert abs(g.loop_meas[g.n_true:, :2].std() - 0.3) < 0.05


def test_named_config_sizes():
    for name, (n, m, k) in {"intel": (1228, 256, 100), "m3500": (3500, 1954, 1000)}.items():
        g, cfg = synth.make_config(name)
        assert (g.n_poses, g.n_true, g.n_loops - g.n_true) == (n, m, k)
        assert cfg["fast_reject_th"] == 6.251
