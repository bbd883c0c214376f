"""The WHOLE M3500-shaped sequential stream (2954 candidates, clusters up to K = 412 loops: BASELINE.json configs[1] run the way the
reference's own tester runs it, src/simulation.cpp:34-47) against the committed oracle fixture of the same stream
(tests/golden/stream_se2_m3500_full.npz, scripts/oracle_full_stream.py: 6 minutes of one CPU core, same termination rule as the GPU
default). Runs last in the GPU suite (file name): ~40 s of GPU time.

Evidence already archived for this pair of runs: the GPU stream (profiles/r02_stream_m3500.json) and the oracle
(profiles/r02_oracle_stream_m3500_full.json) agree on accepted = 411, true / false positives = 379 / 32 and K_max = 412."""
import os

import numpy as np
import pytest

from ipc_b200 import synth
from tests.golden_util import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def test_m3500_full_stream_inlier_set_matches_oracle_fixture(gpu_lib):
    z = np.load(os.path.join(GOLDEN, "stream_se2_m3500_full.npz"))
    g, cfg = synth.make_config("m3500")
    o = g.time_order()
    assert np.array_equal(o, z["order"])
    ipc = gpu_lib.IPC.from_graph(g, cfg, candidates=False)
    acc, info = ipc.agreementCheckStream(g.loop_from[o], g.loop_to[o], g.loop_meas[o], g.loop_info[o])
    acc = np.asarray(acc, dtype=bool)
    assert int(acc.sum()) == int(z["accept"].sum()) == 411
    assert np.array_equal(acc, z["accept"])                                   # bit-exact inlier / outlier membership
    assert np.array_equal(info["n_loops"], z["n_cluster"] + 1) and int(info["n_loops"].max()) == 412
    rel = rel_err(info["max_chi2"], z["max_chi2"])
    assert np.median(rel) < 1e-6 and (rel > 1e-4).mean() <= 0.01              # the decision quantity, candidate by candidate
    ipc.close()


def test_cpp_class_runs_on_the_device(gpu_lib, tmp_path):
    """include/ipc_b200.hpp end to end on the GPU: the check program of tests/test_cpp_wrapper.py constructs an IPC2D on a small chain,
    calls agreementCheck and reads the consensus set back; any C-ABI error would surface as ipc_b200::Error -> WRAPPER_FAIL."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "wrapper_check")
    lib_dir = os.path.join(root, "ipc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "wrapper_check.cpp"),
                           "-L" + lib_dir, "-lipc_b200", "-Wl,-rpath," + lib_dir, "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "WRAPPER_OK" in p.stdout and "device run: agreementCheck" in p.stdout, p.stdout + p.stderr
