"""Pin the oracle's SE(2)/SE(3) algebra (g2o semantics): known answers, analytic
Jacobians against central finite differences under the vertex oplus, group identities."""
import math

import numpy as np
import pytest


def _rand_pose(rng, dim, scale=2.0):
    if dim == 2:
        return np.array([rng.normal() * scale, rng.normal() * scale, rng.uniform(-math.pi, math.pi)])
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    if q[3] < 0:
        q = -q
    return np.concatenate([rng.normal(size=3) * scale, q])


def test_se2_known_answers(oracle_lib):
    po = oracle_lib
    # compose: (1, 2, pi/2) * (1, 0, pi/2) = (1, 3, -pi)  (theta normalised into [-pi, pi))
    r = po.compose(2, [1, 2, math.pi / 2], [1, 0, math.pi / 2])
    assert np.allclose(r[:2], [1, 3])
    assert r[2] == pytest.approx(-math.pi)
    inv = po.inverse(2, [1, 2, math.pi / 2])
    assert np.allclose(inv, [-2, 1, -math.pi / 2])
    # VertexSE2::oplusImpl adds the translation in the GLOBAL frame
    assert np.allclose(po.oplus(2, [1, 2, 0.5], [0.1, 0.2, 0.3]), [1.1, 2.2, 0.8])
    # error of a perfectly satisfied edge is zero
    e, _, _ = po.edge_eval(2, po.compose(2, po.inverse(2, [1, 2, 0.3]), [2, 3, 1.0]), [1, 2, 0.3], [2, 3, 1.0])
    assert np.allclose(e, 0, atol=1e-14)


def test_se3_known_answers(oracle_lib):
    po = oracle_lib
    s = math.sin(math.pi / 4)
    a = [1, 0, 0, 0, 0, s, s]            # 90 deg about z
    r = po.compose(3, a, [1, 0, 0, 0, 0, 0, 1])
    assert np.allclose(r, [1, 1, 0, 0, 0, s, s])
    ident = po.compose(3, a, po.inverse(3, a))
    assert np.allclose(ident, [0, 0, 0, 0, 0, 0, 1], atol=1e-15)
    # oplus: right multiplication by (t, compact quaternion)
    r = po.oplus(3, a, [1, 0, 0, 0, 0, 0])
    assert np.allclose(r, [1, 1, 0, 0, 0, s, s])


@pytest.mark.parametrize("dim", [2, 3])
def test_jacobians_match_finite_differences(oracle_lib, dim):
    po = oracle_lib
    rng = np.random.default_rng(7 + dim)
    d = 3 if dim == 2 else 6
    for _ in range(25):
        xi, xj = _rand_pose(rng, dim), _rand_pose(rng, dim)
        z = po.compose(dim, po.compose(dim, po.inverse(dim, xi), xj), _rand_pose(rng, dim, 0.05) if dim == 2 else
                       np.concatenate([rng.normal(size=3) * 0.05, [0.02, -0.01, 0.03, 1.0]]))
        e, Ji, Jj = po.edge_eval(dim, z, xi, xj)
        h = 1e-6
        Ni, Nj = np.zeros((d, d)), np.zeros((d, d))
        for c in range(d):
            u = np.zeros(d)
            u[c] = h
            ep, _, _ = po.edge_eval(dim, z, po.oplus(dim, xi, u), xj)
            em, _, _ = po.edge_eval(dim, z, po.oplus(dim, xi, -u), xj)
            Ni[:, c] = (ep - em) / (2 * h)
            ep, _, _ = po.edge_eval(dim, z, xi, po.oplus(dim, xj, u))
            em, _, _ = po.edge_eval(dim, z, xi, po.oplus(dim, xj, -u))
            Nj[:, c] = (ep - em) / (2 * h)
        assert np.allclose(Ji, Ni, atol=2e-7, rtol=1e-6)
        assert np.allclose(Jj, Nj, atol=2e-7, rtol=1e-6)


def test_normalize_theta_range(oracle_lib):
    po = oracle_lib
    for t in [-10.0, -math.pi, -3.0, 0.0, 3.0, math.pi, 7.0, 100.0]:
        r = po.compose(2, [0, 0, t], [0, 0, 0])[2]
        assert -math.pi <= r < math.pi
        assert math.isclose(math.cos(r), math.cos(t), abs_tol=1e-12) and math.isclose(math.sin(r), math.sin(t), abs_tol=1e-12)
