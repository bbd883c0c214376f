"""Sequential stream (IPC::agreementCheck per candidate, src/simulation.cpp:34-47) on the GPU vs the CPU oracle.
Usage: python scripts/stream_bench.py intel 1.0 [--oracle]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ipc_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "intel"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with_oracle = "--oracle" in sys.argv
g, cfg = synth.make_config(name, scale=scale)
order = g.time_order()
for a in sys.argv:
    if a.startswith("--limit="):
        order = order[: int(a.split("=")[1])]
t = time.perf_counter()
ipc = api.IPC.from_graph(g, cfg, candidates=False)
create_s = time.perf_counter() - t
acc = np.zeros(len(order), dtype=bool); mx = np.zeros(len(order)); K = np.zeros(len(order), dtype=int); L = np.zeros(len(order), dtype=int)
ev = np.zeros(len(order), dtype=int)
depth = 0
for a in sys.argv:
    if a.startswith("--depth="):
        depth = int(a.split("=")[1])
t = time.perf_counter()
if depth:            # the whole candidate loop in one call (speculative side-by-side solves, sequential semantics)
    ipc.set_option("stream_depth", depth)
    acc, inf = ipc.agreementCheckStream(g.loop_from[order], g.loop_to[order], g.loop_meas[order], g.loop_info[order])
    mx, K, L, ev = inf["max_chi2"], inf["n_loops"], inf["window_len"], inf["evals"]
else:
    for k, l in enumerate(order):
        ok, ci = ipc.agreementCheck((g.loop_from[l], g.loop_to[l], g.loop_meas[l], g.loop_info[l]))
        acc[k] = ok; mx[k] = ci.max_chi2; K[k] = ci.n_loops; L[k] = ci.window_len; ev[k] = ci.evals
dt = time.perf_counter() - t
truth = order < g.n_true
tp = int((acc & truth).sum()); fp = int((acc & ~truth).sum()); fn = int((~acc & truth).sum())
out = {"stream_depth": depth, "config": name, "scale": scale, "create_s": create_s, "n_poses": g.n_poses, "candidates": len(order), "gpu_stream_s": dt, "gpu_checks_per_s": len(order) / dt,
       "accepted": int(acc.sum()), "precision": tp / max(1, tp + fp), "recall": tp / max(1, tp + fn), "K_median": float(np.median(K)), "K_max": int(K.max()),
       "L_median": float(np.median(L)), "evals_mean": float(ev.mean())}
out["profile"] = ipc.stream_profile()
if with_oracle:
    from oracle import pyoracle as po
    t = time.perf_counter(); oacc, orep = po.OracleIPC(g, cfg, noise_exit=True).run_stream(order); odt = time.perf_counter() - t
    rel = np.abs(mx - orep["max_chi2"]) / np.maximum(np.abs(orep["max_chi2"]), 1e-9)
    out.update({"oracle_stream_s": odt, "oracle_checks_per_s": len(order) / odt, "verdict_mismatches": int((acc != oacc).sum()), "max_rel_chi2_err": float(rel.max()),
                "oracle_cores": 1})
print(json.dumps(out))
