"""Host logic of the Monte-Carlo protocol command (scripts/montecarlo.py; bash/ipc_experiments_2D.sh:3-41 of the reference): directory
layout, the yq overrides, the ten-runs-side-by-side loop and the .PR collection — driven with a stub tester, no GPU."""
import importlib.util
import json
import os
import stat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("montecarlo", os.path.join(ROOT, "scripts", "montecarlo.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


STUB = """#!/usr/bin/env python
import sys
cfg = dict(l.strip().split(": ", 1) for l in open(sys.argv[sys.argv.index("-c") + 1]) if ": " in l)
out = cfg["output"].strip('"')
n = sum(1 for l in open(cfg["dataset"].strip('"')) if l.startswith("EDGE_SE2") and abs(int(l.split()[2]) - int(l.split()[1])) != 1)
assert cfg["s_factor"] == "10.0" and cfg["fast_reject_th"] == "10.64" and cfg["slow_reject_th"] == "10.64" and cfg["use_recovery"] == "true"
assert cfg["k_buddies"] == "2" and cfg["use_best_k_buddies"] == "false"
open(out, "w").write("0 0 0\\n")
open(out[:-3] + "PR", "w").write("1 %g\\n%g %g\\n" % (int(cfg["canonic_inliers"]) / n, 0.5, 0.5 / n))
"""


def test_protocol_layout_and_summary(tmp_path):
    mc = _load()
    stub = tmp_path / "stub_tester.py"
    stub.write_text(STUB)
    os.chmod(stub, os.stat(stub).st_mode | stat.S_IEXEC)
    s = mc.main(["--dataset", "intel", "--scale", "0.1", "--outliers", "5,10", "--runs", "3", "--jobs", "2", "--workdir", str(tmp_path / "mc"),
                 "--tester", str(stub), "--tester-args", "", "--date", "010101", "--opt", "T"])
    root = tmp_path / "mc" / "INTEL"
    for out in (5, 10):
        for run in ("00", "01", "02"):
            assert (root / "SPOILED_DATA" / str(out) / f"{run}.g2o").exists()
            assert (root / "EXP" / "010101" / "T" / str(out) / f"{run}.PR").exists()
    assert not list(root.glob("*.yaml"))                        # `rm ./*.yaml`
    true_loops = s["true_loops"]
    assert s["levels"]["5"]["runs"] == 3 and s["levels"]["10"]["candidates"] == true_loops + 10
    assert abs(s["levels"]["10"]["recall_mean"] - true_loops / (true_loops + 10)) < 1e-6
    assert json.load(open(root / "summary.json"))["levels"]["5"]["precision_mean"] == 1.0
    # the spoiled graphs differ between runs and keep the true loops first
    a = open(root / "SPOILED_DATA" / "5" / "00.g2o").read()
    b = open(root / "SPOILED_DATA" / "5" / "01.g2o").read()
    assert a != b and a.split("\n")[1] == b.split("\n")[1]
