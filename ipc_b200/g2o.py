"""g2o text-format graph files + the flat YAML config of the testers (host-side helpers for tests, bench and the CLI).

Format (g2o tag 20201223_git): ``VERTEX_SE2 id x y th``, ``EDGE_SE2 i j dx dy dth I11 I12 I13 I22 I23 I33``,
``VERTEX_SE3:QUAT id x y z qx qy qz qw``, ``EDGE_SE3:QUAT i j x y z qx qy qz qw`` + 21 upper-triangular information entries.
Loop edges are written true loops first, outliers last (what scripts/generateDataset.py:188-246 produces and what
src/simulation.cpp:24-25 assumes)."""
from __future__ import annotations

import numpy as np

from .synth import Graph


def _upper(info):
    d = info.shape[0]
    return [info[r, c] for r in range(d) for c in range(r, d)]


def write_g2o(graph: Graph, path: str, poses=None) -> None:
    two = graph.dim == 2
    vt, et = ("VERTEX_SE2", "EDGE_SE2") if two else ("VERTEX_SE3:QUAT", "EDGE_SE3:QUAT")
    n = graph.n_poses
    if poses is None:
        poses = graph.gt if graph.gt is not None else np.zeros((n, 3 if two else 7))
    with open(path, "w") as f:
        f.write(f"# synthetic graph: {graph.meta}\n")
        for i in range(n):
            f.write(vt + f" {i} " + " ".join(repr(float(x)) for x in poses[i]) + "\n")
        for j in range(n - 1):
            f.write(et + f" {j} {j + 1} " + " ".join(repr(float(x)) for x in list(graph.odom_meas[j]) + _upper(graph.odom_info[j])) + "\n")
        for l in range(graph.n_loops):
            f.write(et + f" {int(graph.loop_from[l])} {int(graph.loop_to[l])} " +
                    " ".join(repr(float(x)) for x in list(graph.loop_meas[l]) + _upper(graph.loop_info[l])) + "\n")


def read_g2o(path: str, dim: int, n_true: int | None = None) -> Graph:
    two = dim == 2
    vt, et = ("VERTEX_SE2", "EDGE_SE2") if two else ("VERTEX_SE3:QUAT", "EDGE_SE3:QUAT")
    mw, d = (3, 3) if two else (7, 6)
    verts, odom, loops = {}, {}, []
    for line in open(path):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] == vt:
            verts[int(t[1])] = [float(x) for x in t[2:2 + mw]]
        elif t[0] == et:
            i, j = int(t[1]), int(t[2])
            vals = [float(x) for x in t[3:]]
            meas, up = vals[:mw], vals[mw:mw + d * (d + 1) // 2]
            info = np.zeros((d, d))
            q = 0
            for r in range(d):
                for c in range(r, d):
                    info[r, c] = info[c, r] = up[q]
                    q += 1
            if abs(j - i) == 1:
                if j != i + 1:
                    raise ValueError("odometry edge not oriented i -> i+1")
                odom[i] = (meas, info)
            else:
                loops.append((i, j, meas, info))
    n = max(max(verts) + 1 if verts else 0, max(odom) + 2 if odom else 0)
    om = np.array([odom[j][0] for j in range(n - 1)])
    oi = np.array([odom[j][1] for j in range(n - 1)])
    gt = np.array([verts[i] for i in range(n)]) if len(verts) == n else None
    return Graph(dim, n, om, oi, np.array([l[0] for l in loops], dtype=np.int32), np.array([l[1] for l in loops], dtype=np.int32),
                 np.array([l[2] for l in loops]).reshape(len(loops), mw), np.array([l[3] for l in loops]).reshape(len(loops), d, d),
                 len(loops) if n_true is None else n_true, gt, {"source": path})


def write_trajectory(poses, path: str) -> None:
    with open(path, "w") as f:
        for p in poses:
            f.write(" ".join(repr(float(x)) for x in p) + "\n")


def write_config(path: str, name: str, dataset: str, ground_truth: str, output: str, canonic_inliers: int, cfg: dict, complete: bool = True) -> None:
    """The 14 keys src/utils.cpp:321-334 reads. ``complete=False`` writes only what the shipped cfg/*.yaml contain (no s_factor etc.)."""
    lines = [f"name: \"{name}\"", f"dataset: \"{dataset}\"", f"ground_truth: \"{ground_truth}\"", f"output: \"{output}\"", "visualize: false",
             f"canonic_inliers: {canonic_inliers}", f"fast_reject_th: {cfg['fast_reject_th']}", f"fast_reject_iter_base: {cfg['fast_reject_iter_base']}",
             f"slow_reject_th: {cfg['slow_reject_th']}", f"slow_reject_iter_base: {cfg['slow_reject_iter_base']}"]
    if complete:
        lines += [f"s_factor: {cfg['s_factor']}", "use_best_k_buddies: false", "k_buddies: 2", "use_recovery: false"]
    open(path, "w").write("\n".join(lines) + "\n")
