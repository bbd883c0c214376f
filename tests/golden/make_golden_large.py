"""Large-K stream fixture: the first N time-ordered candidates of the FULL-SIZE M3500-shaped stream (BASELINE.json configs[1]) through
the CPU oracle (g2o's full retry rule), reaching clusters of K >= 250 loops. Stores only the per-candidate outputs and the final
optimisation of the resulting consensus set (the graph is regenerated from its seed by synth.make_config("m3500")).
Run:  python tests/golden/make_golden_large.py [N]      (minutes of CPU)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ipc_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1700
    g, cfg = synth.make_config("m3500")
    order = g.time_order()[:n]
    orc = po.OracleIPC(g, cfg)
    t = time.time()
    acc, rep = orc.run_stream(order)
    print("stream", time.time() - t, "s; accepted", int(acc.sum()), "K max", int(rep["n_cluster"].max()) + 1)
    poses_stream = orc.poses()
    t = time.time()
    chi2, it = orc.final_optimize(1000)
    print("final optimisation", time.time() - t, "s", chi2, it)
    np.savez_compressed(os.path.join(HERE, "stream_se2_m3500_prefix.npz"), n=n, order=order, accept=acc, max_chi2=rep["max_chi2"], cand_chi2=rep["cand_chi2"],
                        n_cluster=rep["n_cluster"], lo=rep["lo"], hi=rep["hi"], consensus=orc.consensus(), poses_stream=poses_stream,
                        final_chi2=chi2, final_iterations=it, poses_final=orc.poses())
