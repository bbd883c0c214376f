"""Second, independent CPU restatement (numpy, dense) of the IPC check — TEST INFRASTRUCTURE ONLY.

Nothing in the reference pins results at the g2o boundary (parity unpinned), so the
C++ oracle (ipc_oracle.hpp) is cross-checked against this deliberately different implementation:
dense normal equations solved with numpy, SE(3) handled with 4x4 homogeneous matrices and Jacobians
taken by central finite differences of the error under the vertex ``oplus`` (instead of the analytic
quaternion formulas), Python control flow for Dogleg and for src/consensus.cpp. Pure-Python loops:
small cases only.

Follows: src/consensus.cpp:42-75,123-171; src/consensus_utils.cpp:6-22,29-43,60-71,98-116,123-130;
g2o Dogleg / SparseOptimizer::optimize semantics per SURVEY.md Appendix A.1-A.6.
"""
from __future__ import annotations

import math

import numpy as np


def _norm_theta(t):
    if -math.pi <= t < math.pi:
        return t
    t = t - math.floor(t / (2 * math.pi)) * 2 * math.pi
    if t >= math.pi:
        t -= 2 * math.pi
    if t < -math.pi:
        t += 2 * math.pi
    return t


# ---- SE2 as 3x3 homogeneous matrices + explicit angle ------------------------------------------
class SE2Geo:
    d = 3

    @staticmethod
    def from_flat(m):
        return np.array([m[0], m[1], _norm_theta(m[2])], dtype=np.float64)

    @staticmethod
    def compose(a, b):
        c, s = math.cos(a[2]), math.sin(a[2])
        return np.array([a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], _norm_theta(a[2] + b[2])])

    @staticmethod
    def inverse(a):
        th = _norm_theta(-a[2])
        c, s = math.cos(th), math.sin(th)
        return np.array([c * -a[0] - s * -a[1], s * -a[0] + c * -a[1], th])

    @staticmethod
    def oplus(x, u):
        return np.array([x[0] + u[0], x[1] + u[1], _norm_theta(x[2] + u[2])])

    @classmethod
    def error(cls, z, xi, xj):
        return cls.compose(cls.inverse(z), cls.compose(cls.inverse(xi), xj))

    @classmethod
    def jac(cls, z, xi, xj):
        # analytic, derived independently: e_t = Rz^T (Ri^T (tj - ti) - tz), e_th = thj - thi - thz
        ci, si = math.cos(xi[2]), math.sin(xi[2])
        cz, sz = math.cos(z[2]), math.sin(z[2])
        RiT = np.array([[ci, si], [-si, ci]])
        RzT = np.array([[cz, sz], [-sz, cz]])
        dt = xj[:2] - xi[:2]
        dRiT = np.array([[-si, ci], [-ci, -si]])     # d(Ri^T)/dth
        Ji = np.zeros((3, 3)); Jj = np.zeros((3, 3))
        Ji[:2, :2] = -RzT @ RiT
        Ji[:2, 2] = RzT @ (dRiT @ dt)
        Ji[2, 2] = -1
        Jj[:2, :2] = RzT @ RiT
        Jj[2, 2] = 1
        return Ji, Jj


# ---- SE3 as 4x4 matrices ------------------------------------------------------------------------
def _R_from_q(x, y, z, w):
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _q_from_R(R):
    tr = np.trace(R)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    q = q / np.linalg.norm(q)
    return -q if q[3] < 0 else q


class SE3Geo:
    d = 6

    @staticmethod
    def from_flat(m):
        q = np.asarray(m[3:7], dtype=np.float64)
        q = q / np.linalg.norm(q)
        if q[3] < 0:
            q = -q
        T = np.eye(4)
        T[:3, :3] = _R_from_q(*q)
        T[:3, 3] = m[:3]
        return T

    @staticmethod
    def compose(a, b):
        return a @ b

    @staticmethod
    def inverse(a):
        T = np.eye(4)
        T[:3, :3] = a[:3, :3].T
        T[:3, 3] = -a[:3, :3].T @ a[:3, 3]
        return T

    @staticmethod
    def oplus(x, u):
        inc = np.eye(4)
        inc[:3, 3] = u[:3]
        w = 1.0 - float(u[3:] @ u[3:])
        if w >= 0:
            inc[:3, :3] = _R_from_q(u[3], u[4], u[5], math.sqrt(w))
        return x @ inc

    @classmethod
    def error(cls, z, xi, xj):
        E = cls.inverse(z) @ cls.inverse(xi) @ xj
        q = _q_from_R(E[:3, :3])
        return np.concatenate([E[:3, 3], q[:3]])

    @classmethod
    def jac(cls, z, xi, xj, h=1e-6):
        Ji = np.zeros((6, 6)); Jj = np.zeros((6, 6))
        for c in range(6):
            u = np.zeros(6); u[c] = h
            Ji[:, c] = (cls.error(z, cls.oplus(xi, u), xj) - cls.error(z, cls.oplus(xi, -u), xj)) / (2 * h)
            Jj[:, c] = (cls.error(z, xi, cls.oplus(xj, u)) - cls.error(z, xi, cls.oplus(xj, -u))) / (2 * h)
        return Ji, Jj


# ---- dense Dogleg --------------------------------------------------------------
def optimize(G, est, edges, free_ids, iterations):
    """edges: list of (from, to, z, info). est: list of poses (modified in place). Returns (iters, evals)."""
    d = G.d
    idx = {v: k for k, v in enumerate(free_ids)}
    n = len(free_ids) * d

    def chi_all():
        return [float(G.error(z, est[a], est[b]) @ W @ G.error(z, est[a], est[b])) for (a, b, z, W) in edges]

    delta, lam, was_pd = 1e4, 1e-7, True
    iters = evals = 0
    ok = True
    it = 0
    while it < iterations and ok:
        cur = sum(chi_all())
        H = np.zeros((n, n)); b = np.zeros(n)
        for (a, bb, z, W) in edges:
            e = G.error(z, est[a], est[bb])
            Ji, Jj = G.jac(z, est[a], est[bb])
            ia, ib = idx.get(a), idx.get(bb)
            if ia is not None:
                H[ia * d:(ia + 1) * d, ia * d:(ia + 1) * d] += Ji.T @ W @ Ji
                b[ia * d:(ia + 1) * d] -= Ji.T @ W @ e
            if ib is not None:
                H[ib * d:(ib + 1) * d, ib * d:(ib + 1) * d] += Jj.T @ W @ Jj
                b[ib * d:(ib + 1) * d] -= Jj.T @ W @ e
            if ia is not None and ib is not None:
                Hij = Ji.T @ W @ Jj
                H[ia * d:(ia + 1) * d, ib * d:(ib + 1) * d] += Hij
                H[ib * d:(ib + 1) * d, ia * d:(ia + 1) * d] += Hij.T
        alpha = (b @ b) / (b @ H @ b) if (b @ H @ b) != 0 else float("nan")
        hsd = alpha * b
        hsd_n = np.linalg.norm(hsd)
        hgn = None
        good = False
        tries = 0
        while True:
            tries += 1
            if hgn is None:
                solved = False
                while not solved:
                    Hd = H + (0 if was_pd else lam) * np.eye(n)
                    try:
                        L = np.linalg.cholesky(Hd)
                        hgn = np.linalg.solve(L.T, np.linalg.solve(L, b))
                        solved = True
                    except np.linalg.LinAlgError:
                        solved = False
                    was_pd = was_pd and solved
                    if not was_pd:
                        if solved:
                            lam = max(1e-12, lam / 5.0)
                        else:
                            lam *= 10
                            if lam > 1e3:
                                return iters, evals
                hgn_n = np.linalg.norm(hgn)
            if hgn_n < delta:
                hdl = hgn
            elif hsd_n > delta:
                hdl = delta / hsd_n * hsd
            else:
                aux = hgn - hsd
                c = hsd @ aux
                bma = aux @ aux
                if c <= 0:
                    beta = (-c + math.sqrt(c * c + bma * (delta * delta - hsd @ hsd))) / bma
                else:
                    beta = (delta * delta - hsd @ hsd) / (c + math.sqrt(c * c + bma * (delta * delta - hsd @ hsd)))
                hdl = hsd + beta * aux
            lin = -(hdl @ H @ hdl) + 2 * (b @ hdl)
            backup = [est[v].copy() for v in free_ids]
            for k, v in enumerate(free_ids):
                est[v] = G.oplus(est[v], hdl[k * d:(k + 1) * d])
            new = sum(chi_all())
            evals += 1
            if abs(lin) < 1e-12:
                lin = 1e-12
            rho = (cur - new) / lin
            if rho > 0:
                good = True
            else:
                for k, v in enumerate(free_ids):
                    est[v] = backup[k]
            if rho > 0.75:
                delta = max(delta, 3 * np.linalg.norm(hdl))
            elif rho < 0.25:
                delta *= 0.5
            if good or tries >= 100:
                break
        iters += 1
        it += 1
        ok = good and tries < 100
    return iters, evals


# ---- IPC (src/consensus.cpp) ---------------------------------------------------------------------
class NumpyIPC:
    def __init__(self, graph, cfg):
        self.G = SE2Geo if graph.dim == 2 else SE3Geo
        G = self.G
        self.cfg = cfg
        self.n = graph.n_poses
        self.odom = [(j, j + 1, G.from_flat(graph.odom_meas[j]), graph.odom_info[j] * cfg["s_factor"]) for j in range(self.n - 1)]
        self.est = [None] * self.n
        self.est[0] = G.from_flat([0, 0, 0] if graph.dim == 2 else [0, 0, 0, 0, 0, 0, 1])
        for i in range(1, self.n):
            self.est[i] = G.compose(self.est[i - 1], self.odom[i - 1][2])
        self.cns = []

    def subgraph(self, a, b):
        lo, hi = min(a, b), max(a, b)
        inc = [False] * len(self.cns)
        members = []
        found = True
        while found:
            found = False
            for k, (f, t, _, _) in enumerate(self.cns):
                if inc[k]:
                    continue
                t0, t1 = min(f, t), max(f, t)
                if min(t1, hi) - max(t0, lo) <= 0:
                    continue
                lo, hi = min(lo, t0), max(hi, t1)
                inc[k] = True
                found = True
                members.append(k)
        return lo, hi, members

    def agreement_check(self, frm, to, meas, info):
        G = self.G
        cand = (int(frm), int(to), G.from_flat(meas), np.asarray(info, dtype=np.float64))
        lo, hi, members = self.subgraph(cand[0], cand[1])
        slow = len(members) > 0
        th = self.cfg["slow_reject_th"] if slow else self.cfg["fast_reject_th"]
        ib = self.cfg["slow_reject_iter_base"] if slow else self.cfg["fast_reject_iter_base"]
        eset = [self.odom[j] for j in range(lo, hi)] + [self.cns[m] for m in members] + [cand]
        snapshot = [p.copy() for p in self.est]
        iters = ib * 5 if len(eset) > 100 else ib
        optimize(G, self.est, eset, list(range(lo + 1, hi + 1)), iters)
        chis = [float(G.error(z, self.est[a], self.est[b]) @ W @ G.error(z, self.est[a], self.est[b])) for (a, b, z, W) in eset]
        ok = all(c <= th for c in chis)
        if not ok:
            self.est = snapshot
            return False, max(chis), chis[-1]
        self.cns.append(cand)
        for i in range(hi + 1, self.n):
            self.est[i] = G.compose(self.est[i - 1], self.odom[i - 1][2])
        return True, max(chis), chis[-1]
