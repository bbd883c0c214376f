"""Oracle vs the committed golden fixtures (tests/golden/make_golden.py): guards the oracle itself."""
import os

import numpy as np
import pytest

from ipc_b200 import api
from tests.golden_util import load, rel_err


@pytest.mark.parametrize("name", ["pairs_se2_intel.npz", "pairs_se2_m3500.npz", "pairs_se3_sphere.npz"])
def test_oracle_reproduces_pair_fixtures(oracle_lib, name):
    z, g, cfg = load(name)
    ptr, idx = api.checks_to_csr(z["member"], z["cand"])
    acc, rep = oracle_lib.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count())
    assert np.array_equal(acc, z["accept"])
    assert rel_err(rep["max_chi2"], z["max_chi2"]).max() < 1e-9
    assert np.array_equal(rep["lo"], z["lo"]) and np.array_equal(rep["hi"], z["hi"])


@pytest.mark.parametrize("name", ["stream_se2_intel.npz", "stream_se3_sphere.npz"])
def test_oracle_reproduces_stream_fixtures(oracle_lib, name):
    z, g, cfg = load(name)
    orc = oracle_lib.OracleIPC(g, cfg)
    acc, rep = orc.run_stream(z["order"])
    assert np.array_equal(acc, z["accept"])
    assert rel_err(rep["max_chi2"], z["max_chi2"]).max() < 1e-9
    assert np.array_equal(orc.consensus(), z["consensus"])
    assert np.allclose(orc.poses(), z["poses"], atol=1e-12)
