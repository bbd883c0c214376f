"""Host side of the drop-in CLI (cli/ipc_host.hpp): flat YAML config (src/utils.cpp:316-337), g2o graph files
(src/utils.cpp:95-126, 172-189). No compute is called on the CPU box (--parse-only)."""
import os
import subprocess

import numpy as np
import pytest

from ipc_b200 import g2o, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI2 = os.path.join(ROOT, "cli", "ipc_tester_2D")
CLI3 = os.path.join(ROOT, "cli", "ipc_tester_3D")


@pytest.fixture(scope="module")
def clis():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cli")])
    return CLI2, CLI3


def _write_case(tmp_path, name, scale, complete=True):
    g, cfg = synth.make_config(name, scale=scale)
    ds, gt, out, yml = (str(tmp_path / f) for f in ("graph.g2o", "gt.txt", "res.txt", "cfg.yaml"))
    g2o.write_g2o(g, ds)
    g2o.write_trajectory(g.gt, gt)
    g2o.write_config(yml, name, ds, gt, out, g.n_true, cfg, complete=complete)
    return g, cfg, yml


def test_g2o_roundtrip(tmp_path):
    for name, dim in (("intel", 2), ("sphere", 3)):
        g, _ = synth.make_config(name, scale=0.1)
        p = str(tmp_path / f"{name}.g2o")
        g2o.write_g2o(g, p)
        r = g2o.read_g2o(p, dim, n_true=g.n_true)
        assert r.n_poses == g.n_poses and r.n_loops == g.n_loops
        assert np.array_equal(r.loop_from, g.loop_from) and np.array_equal(r.loop_to, g.loop_to)
        assert np.allclose(r.odom_meas, g.odom_meas, rtol=0, atol=0) and np.allclose(r.loop_info, g.loop_info, rtol=0, atol=0)


def test_cli_parses_config_and_graph(clis, tmp_path):
    g, cfg, yml = _write_case(tmp_path, "intel", 0.1)
    out = subprocess.run([clis[0], "-c", yml, "--parse-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert f"{g.n_poses} poses, {g.n_poses - 1} odometry edges, {g.n_loops} loop candidates ({g.n_true} canonic inliers)" in out.stdout
    assert "s_factor 10 fast 6.251/50 slow 11.345/100" in out.stdout


def test_cli_3d_parses(clis, tmp_path):
    g, cfg, yml = _write_case(tmp_path, "sphere", 0.05)
    out = subprocess.run([clis[1], "-c", yml, "--parse-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert f"{g.n_poses} poses" in out.stdout and f"{g.n_loops} loop candidates" in out.stdout


def test_cli_rejects_incomplete_config_like_the_reference(clis, tmp_path):
    """The shipped cfg/*.yaml lack s_factor / use_best_k_buddies / k_buddies / use_recovery; readConfig throws on them."""
    _, _, yml = _write_case(tmp_path, "intel", 0.1, complete=False)
    out = subprocess.run([clis[0], "-c", yml, "--parse-only"], capture_output=True, text=True)
    assert out.returncode == 1 and "config key missing: s_factor" in out.stderr


def test_cli_usage_and_missing_files(clis, tmp_path):
    assert subprocess.run([clis[0]], capture_output=True).returncode == 2
    out = subprocess.run([clis[0], "-c", str(tmp_path / "nope.yaml")], capture_output=True, text=True)
    assert out.returncode == 1 and "cannot open config" in out.stderr


def _inv_meas(m):
    """inverse of a relative pose: (x, y, theta) or (t, qx qy qz qw) — independent numpy restatement for the test"""
    if len(m) == 3:
        c, s = np.cos(m[2]), np.sin(m[2])
        th = -m[2]
        th = (th + np.pi) % (2 * np.pi) - np.pi
        return np.array([-(c * m[0] + s * m[1]), -(-s * m[0] + c * m[1]), th])
    q = m[3:] / np.linalg.norm(m[3:])
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return np.concatenate([-R.T @ m[:3], [-x, -y, -z, w]])


def _flip_loops(path_in, path_out, dim, every=3):
    """Rewrite a g2o file with every `every`-th loop edge stored as (to, from) with the inverse measurement — what graph_fixer undoes."""
    et = "EDGE_SE2" if dim == 2 else "EDGE_SE3:QUAT"
    mw = 3 if dim == 2 else 7
    out, n_loop, flipped = [], 0, 0
    for line in open(path_in):
        t = line.split()
        if t and t[0] == et and abs(int(t[2]) - int(t[1])) != 1:
            n_loop += 1
            if n_loop % every == 0:
                m = np.array([float(x) for x in t[3:3 + mw]])
                inv = _inv_meas(m)
                line = " ".join([t[0], t[2], t[1]] + [repr(float(x)) for x in inv] + t[3 + mw:]) + "\n"
                flipped += 1
        out.append(line)
    open(path_out, "w").writelines(out)
    return flipped


@pytest.mark.parametrize("name,dim,scale", [("intel", 2, 0.2), ("sphere", 3, 0.05)])
def test_graph_fixer_reorients_loops(clis, tmp_path, name, dim, scale):
    """examples/graph_fixer.cpp:36-52: loops stored as to < from come back as from < to with the inverse measurement and the
    information untouched; odometry, vertices and the edge order stay as they are; the saved file reloads bit-exactly."""
    g, cfg, yml = _write_case(tmp_path, name, scale)
    ds = str(tmp_path / "graph.g2o")
    flipped_path = str(tmp_path / "flipped.g2o")
    n_flip = _flip_loops(ds, flipped_path, dim)
    assert n_flip > 0
    yml2 = str(tmp_path / "cfg2.yaml")
    g2o.write_config(yml2, name, flipped_path, str(tmp_path / "gt.txt"), str(tmp_path / "res.txt"), g.n_true, cfg)
    fixed = str(tmp_path / "fixed.g2o")
    fixer = os.path.join(ROOT, "cli", "graph_fixer")
    out = subprocess.run([fixer, "-c", yml2, "-o", fixed, "--dim", str(dim)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert f"Tot Vertices = {g.n_poses}" in out.stdout and f"Re-oriented {n_flip} loop edges" in out.stdout
    r = g2o.read_g2o(fixed, dim, n_true=g.n_true)
    assert np.array_equal(r.loop_from, g.loop_from) and np.array_equal(r.loop_to, g.loop_to)      # same order, from < to again
    assert np.array_equal(r.odom_meas, g.odom_meas) and np.array_equal(r.loop_info, g.loop_info)
    if dim == 3:      # q and -q are the same rotation
        sgn = np.sign(np.sum(r.loop_meas[:, 3:] * g.loop_meas[:, 3:], axis=1))[:, None]
        r.loop_meas[:, 3:] *= sgn
        qn = g.loop_meas[:, 3:] / np.linalg.norm(g.loop_meas[:, 3:], axis=1, keepdims=True)
        ref = np.concatenate([g.loop_meas[:, :3], qn], axis=1)
        touched = np.any(r.loop_meas != g.loop_meas, axis=1)
        assert np.allclose(r.loop_meas[touched], ref[touched], rtol=0, atol=1e-12)
        assert np.array_equal(r.loop_meas[~touched], g.loop_meas[~touched])
    else:
        assert np.allclose(r.loop_meas, g.loop_meas, rtol=0, atol=1e-12)
    # a second pass finds nothing to do and reproduces the file
    yml3 = str(tmp_path / "cfg3.yaml")
    g2o.write_config(yml3, name, fixed, str(tmp_path / "gt.txt"), str(tmp_path / "res.txt"), g.n_true, cfg)
    again = str(tmp_path / "again.g2o")
    out = subprocess.run([fixer, "-c", yml3, "-o", again, "--dim", str(dim)], capture_output=True, text=True)
    assert out.returncode == 0 and "Re-oriented 0 loop edges" in out.stdout
    assert open(fixed).read() == open(again).read()


def test_threaded_loader_keeps_file_order(clis, tmp_path):
    """The chunked / threaded tokeniser of the CLI must give the candidates in FILE order: a file big enough to be cut
    into several chunks is parsed and the CLI's counts match; FIX lines and unknown tags are tolerated."""
    g, cfg, yml = _write_case(tmp_path, "m3500", 1.0)
    ds = str(tmp_path / "graph.g2o")
    with open(ds, "a") as f:
        f.write("FIX 0\nSOME_UNKNOWN_TAG 1 2 3\n")
    assert os.path.getsize(ds) > 4 * (1 << 16)
    out = subprocess.run([clis[0], "-c", yml, "--parse-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert f"{g.n_poses} poses, {g.n_poses - 1} odometry edges, {g.n_loops} loop candidates" in out.stdout
    fixed = str(tmp_path / "fixed.g2o")
    out = subprocess.run([os.path.join(ROOT, "cli", "graph_fixer"), "-c", yml, "-o", fixed, "--dim", "2"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    r = g2o.read_g2o(fixed, 2, n_true=g.n_true)
    assert np.array_equal(r.loop_from, np.minimum(g.loop_from, g.loop_to)) and np.array_equal(r.odom_meas, g.odom_meas)
    keep = g.loop_from < g.loop_to
    assert np.array_equal(r.loop_meas[keep], g.loop_meas[keep])
    assert "FIX 0" in open(fixed).read()


def test_cli_matrix_mode_needs_a_gpu_and_says_so(clis, tmp_path):
    """No CPU fallback: on a box without a CUDA device the batch path of the tester fails loudly (exit code 1, the library's message)."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "ipc_b200", "libipc_b200.so"))
    if lib.ipc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    _, _, yml = _write_case(tmp_path, "intel", 0.1)
    out = subprocess.run([clis[0], "-c", yml, "--matrix"], capture_output=True, text=True)
    assert out.returncode == 1 and "CUDA device" in out.stderr
    out = subprocess.run([clis[0], "-c", yml, "--matrix", "--gpus", "2"], capture_output=True, text=True)
    assert out.returncode == 1 and "CUDA device" in out.stderr
