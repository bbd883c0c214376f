"""The WHOLE M3500-shaped sequential stream through the single-threaded CPU oracle (about 6 minutes): aggregate figures for
profiles/r02_oracle_stream_m3500_full.json and the per-candidate outputs as tests/golden/stream_se2_m3500_full.npz.
Run:  python scripts/oracle_full_stream.py > profiles/r02_oracle_stream_m3500_full.json"""
import json, os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ipc_b200 import synth
from oracle import pyoracle as po
g, cfg = synth.make_config("m3500")
order = g.time_order()
t = time.perf_counter(); acc, rep = po.OracleIPC(g, cfg, noise_exit=True).run_stream(order); dt = time.perf_counter() - t
truth = order < g.n_true
tp = int((acc & truth).sum()); fp = int((acc & ~truth).sum()); fn = int((~acc & truth).sum())
out = {"what": "CPU oracle (oracle/ipc_oracle.hpp, one thread, same termination rule as the GPU default) on the WHOLE M3500-shaped sequential stream",
       "candidates": len(order), "oracle_stream_s": dt, "oracle_checks_per_s": len(order) / dt, "accepted": int(acc.sum()), "true_positives": tp, "false_positives": fp,
       "precision": tp / max(1, tp + fp), "recall": tp / max(1, tp + fn), "K_max": int(rep["n_cluster"].max()) + 1,
       "accept_bits_sha256": hashlib.sha256(np.packbits(acc).tobytes()).hexdigest()}
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "stream_se2_m3500_full.npz"), accept=acc, max_chi2=rep["max_chi2"], n_cluster=rep["n_cluster"], order=order)
print(json.dumps(out))
