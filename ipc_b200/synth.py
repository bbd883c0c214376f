"""Synthetic pose-graph generators for the BASELINE.json configs.

No dataset ships with the reference (README.md:26 links a Google-Drive folder) and there is no
network, so the named shapes are re-created with fixed seeds:

  * Manhattan-world SE(2) walks (INTEL / M3500 / City10000 / 50k shapes): unit steps on a grid,
    +-90 degree turns, proximity loop closures, Gaussian noise consistent with the information
    matrix written into the file.
  * Sphere SE(3): ``rings`` x ``per_ring`` poses on a sphere, loops i <-> i - per_ring (2450 for
    50 x 50, cfg/3D/SPHERE_params.yaml:6).
  * Outliers exactly as /root/reference/scripts/generateDataset.py:188-246 draws them (random
    vertex pair, N(0, 0.3) translation, N(0, 10 deg) rotation, information copied from the first
    true loop; the 3D quaternion is written in (w, x, y, z) order into the (qx, qy, qz, qw)
    slots, reproducing the script's quirk at :225,:237-240). numpy's Generator replaces
    ``random`` so the stream differs but the distribution does not.

A graph is a plain dict of numpy arrays (see ``Graph``), which is also what the g2o reader in
``ipc_b200.g2o`` returns.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

__all__ = ["Graph", "manhattan", "sphere", "add_outliers", "CONFIGS", "make_config", "make_clean"]


@dataclass
class Graph:
    dim: int                      # 2 or 3
    n_poses: int
    odom_meas: np.ndarray         # [N-1, 3] (x, y, th) or [N-1, 7] (x y z qx qy qz qw)
    odom_info: np.ndarray         # [N-1, d, d] full symmetric, d = 3 or 6
    loop_from: np.ndarray         # [M] int32 (file order: true loops first, then outliers)
    loop_to: np.ndarray           # [M] int32
    loop_meas: np.ndarray         # [M, 3 | 7]
    loop_info: np.ndarray         # [M, d, d]
    n_true: int                   # canonic_inliers: the first n_true loops are the true ones
    gt: np.ndarray | None = None  # [N, 3 | 7] ground-truth poses
    meta: dict = field(default_factory=dict)

    @property
    def d(self) -> int:
        return 3 if self.dim == 2 else 6

    @property
    def n_loops(self) -> int:
        return int(self.loop_from.shape[0])

    def time_order(self) -> np.ndarray:
        """Candidate order of src/simulation.cpp:26 (cmpTime: by max vertex id). The reference uses
        an unstable std::sort; the contract here is a STABLE sort over file order."""
        mx = np.maximum(self.loop_from, self.loop_to)
        return np.argsort(mx, kind="stable")


# --------------------------------------------------------------------------------------------
# SE(2) helpers
# --------------------------------------------------------------------------------------------
def _wrap(a):
    return (a + math.pi) % (2 * math.pi) - math.pi


def _se2_rel(a, b):
    """a^-1 * b for rows (x, y, th)."""
    c, s = np.cos(a[..., 2]), np.sin(a[..., 2])
    dx, dy = b[..., 0] - a[..., 0], b[..., 1] - a[..., 1]
    return np.stack([c * dx + s * dy, -s * dx + c * dy, _wrap(b[..., 2] - a[..., 2])], axis=-1)


def manhattan(n_poses: int, n_loops: int, seed: int, info_diag=(44.721360, 44.721360, 30.901699),
              noise_scale: float = 1.0, turn_prob: float = 0.3, box: int | None = None,
              min_gap: int = 10, reverse_frac: float = 0.0) -> Graph:
    """Manhattan-world walk with ``n_loops`` proximity loop closures (true loops only)."""
    rng = np.random.default_rng(seed)
    if box is None:
        box = max(8, int(round(math.sqrt(n_poses) * 0.45)))
    dirs = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]])
    pos = np.zeros((n_poses, 2), dtype=np.int64)
    head = np.zeros(n_poses, dtype=np.int64)
    h, p = 0, np.array([0, 0])
    for i in range(1, n_poses):
        if rng.random() < turn_prob:
            h = (h + (1 if rng.random() < 0.5 else -1)) % 4
        q = p + dirs[h]
        tries = 0
        while (abs(q[0]) > box or abs(q[1]) > box) and tries < 8:
            h = (h + (1 if rng.random() < 0.5 else -1)) % 4
            q = p + dirs[h]
            tries += 1
        p = q
        pos[i], head[i] = p, h
    gt = np.column_stack([pos.astype(np.float64), _wrap(head * (math.pi / 2))])
    # proximity loops: later pose j revisits a cell (or a 4-neighbour) seen at i, j - i > min_gap
    cells: dict[tuple[int, int], list[int]] = {}
    cand = []
    for j in range(n_poses):
        key = (int(pos[j, 0]), int(pos[j, 1]))
        for dk in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)):
            lst = cells.get((key[0] + dk[0], key[1] + dk[1]))
            if lst:
                for i in lst[-3:]:
                    if j - i > min_gap:
                        cand.append((i, j))
        cells.setdefault(key, []).append(j)
    cand = sorted(set(cand), key=lambda ij: (ij[1], ij[0]))
    if len(cand) < n_loops:
        raise ValueError(f"walk produced only {len(cand)} loop candidates, need {n_loops}")
    pick = np.sort(rng.choice(len(cand), size=n_loops, replace=False))
    pairs = np.array([cand[k] for k in pick], dtype=np.int32)
    sig = noise_scale / np.sqrt(np.asarray(info_diag))
    odom = _se2_rel(gt[:-1], gt[1:]) + rng.normal(size=(n_poses - 1, 3)) * sig
    odom[:, 2] = _wrap(odom[:, 2])
    lf, lt = pairs[:, 0].copy(), pairs[:, 1].copy()
    rev = rng.random(n_loops) < reverse_frac
    lf[rev], lt[rev] = pairs[rev, 1], pairs[rev, 0]
    lmeas = _se2_rel(gt[lf], gt[lt]) + rng.normal(size=(n_loops, 3)) * sig
    lmeas[:, 2] = _wrap(lmeas[:, 2])
    info = np.diag(np.asarray(info_diag, dtype=np.float64))
    return Graph(2, n_poses, odom, np.broadcast_to(info, (n_poses - 1, 3, 3)).copy(), lf, lt, lmeas,
                 np.broadcast_to(info, (n_loops, 3, 3)).copy(), n_loops, gt,
                 {"kind": "manhattan", "seed": seed, "info_diag": list(info_diag), "noise_scale": noise_scale})


# --------------------------------------------------------------------------------------------
# SE(3) helpers (quaternions stored x, y, z, w like g2o files)
# --------------------------------------------------------------------------------------------
def _qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _qrot(q, v):
    qv = np.concatenate([v, np.zeros(v.shape[:-1] + (1,))], axis=-1)
    return _qmul(_qmul(q, qv), _qconj(q))[..., :3]


def _q_from_R(R):
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def _se3_rel(a, b):
    """a^-1 * b for rows (x y z qx qy qz qw)."""
    qa_inv = _qconj(a[..., 3:7])
    t = _qrot(qa_inv, b[..., :3] - a[..., :3])
    q = _qmul(qa_inv, b[..., 3:7])
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    q = np.where(q[..., 3:4] < 0, -q, q)
    return np.concatenate([t, q], axis=-1)


def _se3_perturb(rel, rng, sig_t, sig_rot):
    n = rel.shape[0]
    out = rel.copy()
    out[:, :3] += rng.normal(size=(n, 3)) * sig_t
    rv = rng.normal(size=(n, 3)) * sig_rot          # rotation vector
    ang = np.linalg.norm(rv, axis=1, keepdims=True)
    half = 0.5 * ang
    dq = np.concatenate([np.where(ang > 1e-12, np.sin(half) / np.maximum(ang, 1e-300), 0.5) * rv, np.cos(half)], axis=1)
    q = _qmul(rel[:, 3:7], dq)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    out[:, 3:7] = np.where(q[:, 3:4] < 0, -q, q)
    return out


def sphere(rings: int, per_ring: int, seed: int, radius: float = 50.0, sig_t: float = 0.01, sig_rot: float = 0.01,
           noise_scale: float = 1.0) -> Graph:
    """g2o create_sphere-style SE(3) graph: a spiral over a sphere, loops to the pose one ring below.
    Information = diag(1/sig_t^2 x3, 4/sig_rot^2 x3) (the MQT quaternion part is half the rotation vector)."""
    rng = np.random.default_rng(seed)
    n = rings * per_ring
    gt = np.zeros((n, 7))
    for r in range(rings):
        phi = -math.pi / 2 + math.pi * (r + 1) / (rings + 1)
        for s in range(per_ring):
            k = r * per_ring + s
            th = 2 * math.pi * s / per_ring
            c = np.array([radius * math.cos(phi) * math.cos(th), radius * math.cos(phi) * math.sin(th), radius * math.sin(phi)])
            # x axis tangent along the ring, z axis pointing outwards
            z = c / np.linalg.norm(c)
            x = np.array([-math.sin(th), math.cos(th), 0.0])
            y = np.cross(z, x)
            R = np.column_stack([x, y, z])
            gt[k, :3] = c
            gt[k, 3:] = _q_from_R(R)
    odom = _se3_perturb(_se3_rel(gt[:-1], gt[1:]), rng, sig_t * noise_scale, sig_rot * noise_scale)
    lt = np.arange(per_ring, n, dtype=np.int32)
    lf = lt - per_ring
    lmeas = _se3_perturb(_se3_rel(gt[lf], gt[lt]), rng, sig_t * noise_scale, sig_rot * noise_scale)
    info = np.diag([1 / sig_t ** 2] * 3 + [4 / sig_rot ** 2] * 3)
    m = lt.shape[0]
    return Graph(3, n, odom, np.broadcast_to(info, (n - 1, 6, 6)).copy(), lf, lt, lmeas,
                 np.broadcast_to(info, (m, 6, 6)).copy(), m, gt,
                 {"kind": "sphere", "seed": seed, "sig_t": sig_t, "sig_rot": sig_rot, "noise_scale": noise_scale})


# --------------------------------------------------------------------------------------------
# outliers, scripts/generateDataset.py:188-246
# --------------------------------------------------------------------------------------------
def _euler_to_quat_wxyz(yaw, pitch, roll):
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    return (cy * cp * cr + sy * sp * sr, cy * cp * sr - sy * sp * cr, cy * sp * cr + sy * cp * sr, sy * cp * cr - cy * sp * sr)


def add_outliers(g: Graph, n_out: int, seed: int, local: bool = False) -> Graph:
    rng = np.random.default_rng(seed)
    N = g.n_poses
    lf, lt, lm = [], [], []
    for _ in range(n_out):
        v1 = v2 = 0
        while v1 == v2:
            v1 = int(rng.integers(0, N - 1))                      # randint(0, poseCount-1-groupSize) inclusive
            v2 = int(rng.integers(v1, min(N - 2, v1 + 20) + 1)) if local else int(rng.integers(0, N - 1))
            if v1 > v2:
                v1, v2 = v2, v1
            if v2 == v1 + 1:
                v2 = v1 + 2
        if g.dim == 2:
            m = [rng.normal(0, 0.3), rng.normal(0, 0.3), rng.normal(0, math.radians(10))]
        else:
            t = [rng.normal(0, 0.3) for _ in range(3)]
            sg = math.radians(10)
            roll, pitch, yaw = rng.normal(0, sg), rng.normal(0, sg), rng.normal(0, sg)
            q0, q1, q2, q3 = _euler_to_quat_wxyz(yaw, pitch, roll)
            m = t + [q0, q1, q2, q3]          # (w,x,y,z) lands in the (qx,qy,qz,qw) slots — reference quirk
        lf.append(v1); lt.append(v2); lm.append(m)
    info = np.broadcast_to(g.loop_info[0], (n_out, g.d, g.d))
    return Graph(g.dim, N, g.odom_meas, g.odom_info,
                 np.concatenate([g.loop_from, np.array(lf, dtype=np.int32)]),
                 np.concatenate([g.loop_to, np.array(lt, dtype=np.int32)]),
                 np.concatenate([g.loop_meas, np.array(lm, dtype=np.float64).reshape(n_out, -1)]),
                 np.concatenate([g.loop_info, info]), g.n_true, g.gt,
                 dict(g.meta, outliers=n_out, outlier_seed=seed))


# --------------------------------------------------------------------------------------------
# the named configs of BASELINE.json
# --------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (builder kwargs, outliers, ipc config) — thresholds / iteration bases from the shipped yaml of the
    # matching dataset, s_factor from bash/ipc_experiments_{2D,3D}.sh:28
    "intel": dict(kind="manhattan", n_poses=1228, n_loops=256, seed=1, info_diag=(11.11, 400.0, 2496.8), outliers=100,
                  cfg=dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)),
    "m3500": dict(kind="manhattan", n_poses=3500, n_loops=1954, seed=2, info_diag=(44.721360, 44.721360, 30.901699), outliers=1000,
                  cfg=dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)),
    "sphere": dict(kind="sphere", rings=50, per_ring=50, seed=3, outliers=2000, noise_scale=0.25,
                   cfg=dict(s_factor=50.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=6.251, slow_reject_iter_base=100)),
    "city10k": dict(kind="manhattan", n_poses=10000, n_loops=10688, seed=4, info_diag=(44.721360, 44.721360, 30.901699), outliers=5000,
                    cfg=dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)),
    "synth50k": dict(kind="manhattan", n_poses=50000, n_loops=5000, seed=5, info_diag=(44.721360, 44.721360, 30.901699), outliers=5000,
                     cfg=dict(s_factor=10.0, fast_reject_th=6.251, fast_reject_iter_base=50, slow_reject_th=11.345, slow_reject_iter_base=100)),
}


def make_clean(name: str, scale: float = 1.0, noise_scale: float | None = None) -> Graph:
    """The named shape WITHOUT outliers (true loops only): what the Monte-Carlo protocol spoils (scripts/montecarlo.py)."""
    c = CONFIGS[name]
    if noise_scale is None:
        # sphere: the file's information (sigma 0.01) is 4x more conservative than the noise actually drawn, as in
        # public datasets; with noise at the stated sigma and s_factor = 50 the reference rejects every loop
        noise_scale = c.get("noise_scale", 1.0)
    if c["kind"] == "manhattan":
        n = max(32, int(round(c["n_poses"] * scale)))
        m = max(4, int(round(c["n_loops"] * scale)))
        g = manhattan(n, m, c["seed"], info_diag=c["info_diag"], noise_scale=noise_scale)
    else:
        rings = max(3, int(round(c["rings"] * math.sqrt(scale))))
        per = max(4, int(round(c["per_ring"] * math.sqrt(scale))))
        g = sphere(rings, per, c["seed"], noise_scale=noise_scale)
    g.meta["config"] = name
    return g


def make_config(name: str, scale: float = 1.0, noise_scale: float | None = None):
    """Build (graph, ipc_cfg) for a named config. ``scale`` < 1 shrinks poses / loops / outliers
    proportionally (parity-test sizes); 1.0 is the BASELINE.json size."""
    c = CONFIGS[name]
    n_out = max(1, int(round(c["outliers"] * scale)))
    g = add_outliers(make_clean(name, scale, noise_scale), n_out, c["seed"] + 1000)
    g.meta["config"] = name
    return g, dict(c["cfg"])
