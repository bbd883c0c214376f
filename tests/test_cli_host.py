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
