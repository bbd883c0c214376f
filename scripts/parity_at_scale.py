"""Full-size parity evidence: a seeded sample of the BASELINE-size check lists, GPU (through the C ABI, with info) against the
CPU oracle with g2o's full retry semantics. Writes one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ipc_b200 import api, synth
from oracle import pyoracle as po

for name, n_s in (("m3500", 16000), ("sphere", 6000), ("intel", 42030)):
    g, cfg = synth.make_config(name)
    mem, cnd = api.pair_checks(g)
    n_s = min(n_s, len(cnd))
    sel = np.sort(np.random.default_rng(11).choice(len(cnd), n_s, replace=False))
    ipc = api.IPC.from_graph(g, cfg)
    t = time.perf_counter(); acc, info = ipc.check_batch(mem[sel], cnd[sel]); tg = time.perf_counter() - t
    ptr, idx = api.checks_to_csr(mem[sel], cnd[sel])
    t = time.perf_counter(); oacc, orep = po.OracleIPC(g, cfg).check_batch(ptr, idx, n_threads=os.cpu_count()); tc = time.perf_counter() - t
    rel = np.abs(info["max_chi2"] - orep["max_chi2"]) / np.maximum(np.abs(orep["max_chi2"]), 1e-9)
    relc = np.abs(info["cand_chi2"] - orep["cand_chi2"]) / np.maximum(np.abs(orep["cand_chi2"]), 1e-9)
    th = np.where(info["n_loops"] == 2, cfg["slow_reject_th"], cfg["fast_reject_th"])
    near = np.abs(orep["max_chi2"] - th) / th
    print(json.dumps({"config": name, "checks": int(n_s), "of": int(len(cnd)), "verdict_mismatches": int((acc != oacc).sum()), "accepted": int(acc.sum()),
                      "max_rel_err_max_chi2": float(rel.max()), "p999_rel_err_max_chi2": float(np.quantile(rel, 0.999)), "n_above_1e-4": int((rel > 1e-4).sum()),
                      "max_rel_err_cand_chi2": float(relc.max()), "closest_to_threshold_rel": float(near.min()), "gpu_s": tg, "oracle_s": tc,
                      "oracle_threads": os.cpu_count(), "mean_window": float(info["window_len"].mean())}), flush=True)
    ipc.close()
