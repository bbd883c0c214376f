"""Outlier injector pinned against the REFERENCE's own output.

tests/golden/spoil/ holds what /root/reference/scripts/generateDataset.py itself wrote for nine option sets
(tests/golden/make_golden_spoil.py ran it, unmodified): sha256 + size of every output file and the appended outlier
lines verbatim. ipc_b200.spoil.spoil_g2o must reproduce each file byte for byte from the same input, seed and options
(same `random` draws in the same order, same str() formatting, same vertex/edge filtering)."""
import hashlib
import json
import os

import numpy as np
import pytest

from ipc_b200 import g2o, spoil, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spoil")
with open(os.path.join(GOLD, "cases.json")) as _f:
    CASES = json.load(_f)


@pytest.mark.parametrize("name", sorted(CASES))
def test_spoil_matches_reference_script_bytes(name, tmp_path):
    c = CASES[name]
    dst = tmp_path / "out.g2o"
    n = spoil.spoil_g2o(os.path.join(GOLD, c["input"]), str(dst), **c["spoil_kwargs"])
    data = dst.read_bytes()
    assert n == c["n_outlier_lines"]
    assert len(data) == c["n_bytes"]
    assert hashlib.sha256(data).hexdigest() == c["sha256"]
    with open(os.path.join(GOLD, f"ref_{name}.tail")) as f:
        tail = f.readlines()
    assert data.decode().splitlines(keepends=True)[-len(tail):] == tail if tail else True


def test_spoiled_file_loads_and_keeps_the_quaternion_slot_quirk(tmp_path):
    """The spoiled 3D file goes through our g2o reader; the (w,x,y,z) tuple sits in the (qx,qy,qz,qw) slots exactly as
    the reference's tester would read it (scripts/generateDataset.py:225,237-240)."""
    c = CASES["3d_default"]
    dst = tmp_path / "s.g2o"
    spoil.spoil_g2o(os.path.join(GOLD, c["input"]), str(dst), **c["spoil_kwargs"])
    clean = g2o.read_g2o(os.path.join(GOLD, c["input"]), 3)
    g = g2o.read_g2o(str(dst), 3, n_true=clean.n_loops)
    assert g.n_loops == clean.n_loops + c["n_outlier_lines"] and g.n_true == clean.n_loops
    with open(os.path.join(GOLD, "ref_3d_default.tail")) as f:
        first = f.readline().split()
    k = clean.n_loops
    assert (int(g.loop_from[k]), int(g.loop_to[k])) == (int(first[1]), int(first[2]))
    np.testing.assert_array_equal(g.loop_meas[k], np.array([float(x) for x in first[3:10]]))
    assert g.loop_meas[k][3] > 0.9           # the slot g2o reads as qx holds w ~ cos(small angle)


def test_in_memory_injector_follows_the_same_rules():
    """synth.add_outliers (numpy Generator stream, used by the seeded bench configs) obeys the index rules the pinned
    file-level injector obeys: 0 <= v1 < v2 <= N-2+1, never neighbours, information of the first true loop."""
    g = synth.add_outliers(synth.make_clean("intel", 0.1), 200, seed=5)
    f, t = g.loop_from[g.n_true:], g.loop_to[g.n_true:]
    assert (f >= 0).all() and (t - f >= 2).all() and (t <= g.n_poses - 1).all()
    np.testing.assert_array_equal(g.loop_info[g.n_true:], np.broadcast_to(g.loop_info[0], (200, 3, 3)))


def test_bad_arguments_fail():
    with pytest.raises(ValueError):
        spoil.spoil_g2o(os.path.join(GOLD, "clean_2d.g2o"), os.devnull, outliers=-1)
    with pytest.raises(ValueError):
        spoil.spoil_g2o(os.path.join(GOLD, "clean_2d.g2o"), os.devnull, outliers=3, information="1,2,3")
