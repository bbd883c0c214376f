"""How the City10000-shaped sequential stream (BASELINE cfg4) grows: the single-threaded CPU oracle on a time-ordered prefix for a
fixed budget — cluster size K and window length L per candidate, calls/s. Feeds DESIGN.md 7 (why the dense dK x dK force system of
the stream solver does not reach this config).  python scripts/oracle_city10k_prefix.py [seconds] > profiles/r02_oracle_city10k_prefix.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ipc_b200 import synth
from oracle import pyoracle as po
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
g, cfg = synth.make_config("city10k")
order = g.time_order()
orc = po.OracleIPC(g, cfg, noise_exit=True)
K, L, acc, ts = [], [], [], []
t0 = time.perf_counter()
for li in order:
    t = time.perf_counter()
    ok, rep = orc.agreement_check(g.loop_from[li], g.loop_to[li], g.loop_meas[li], g.loop_info[li])
    ts.append(time.perf_counter() - t); acc.append(bool(ok)); K.append(int(rep.n_cluster) + 1); L.append(int(rep.hi - rep.lo))
    if time.perf_counter() - t0 > budget:
        break
n = len(K); K = np.array(K); L = np.array(L); ts = np.array(ts)
q = max(1, n // 4)
print(json.dumps({"config": "city10k (10000 poses, 15688 candidates)", "budget_s": budget, "candidates_done": n, "accepted": int(np.sum(acc)), "calls_per_s": n / float(ts.sum()),
                  "K_max": int(K.max()), "K_median_last_quarter": float(np.median(K[-q:])), "L_median_last_quarter": float(np.median(L[-q:])),
                  "s_per_call_last_quarter": float(ts[-q:].mean()), "last_vertex_reached": int(max(g.loop_to[order[n - 1]], g.loop_from[order[n - 1]]))}))
