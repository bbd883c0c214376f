"""Diagnostic: long-window checks (L > 5400) of a config on the GPU against the oracle; prints the chi2 agreement per check."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ipc_b200 import api, synth, sharding
from oracle import pyoracle as po
name = sys.argv[1]; n_checks = int(sys.argv[2])
g, cfg = synth.make_config(name)
mem, cnd = api.pair_checks(g)
L = sharding.window_lengths(g, mem, cnd)
long_ = np.nonzero(L > 5400)[0]
sel = np.sort(np.random.default_rng(4).choice(long_, n_checks, replace=False))
ptr, idx = api.checks_to_csr(mem[sel], cnd[sel])
res = {}
for ne in (True, False):
    oacc, orep = po.OracleIPC(g, cfg, noise_exit=ne).check_batch(ptr, idx, n_threads=os.cpu_count())
    res[ne] = (oacc, orep)
for opts in ({}, {"noise_exit": 0}, {"cta_per_check": 1}):
    ipc = api.IPC.from_graph(g, cfg)
    for k, v in opts.items(): ipc.set_option(k, v)
    acc, info = ipc.check_batch(mem[sel], cnd[sel])
    for ne in (True, False):
        oacc, orep = res[ne]
        rel = np.abs(info["max_chi2"] - orep["max_chi2"]) / np.maximum(np.abs(orep["max_chi2"]), 1e-300)
        w = np.argsort(-rel)[:4]
        print(json.dumps({"lib": os.environ.get("IPC_B200_LIB", "default"), "opts": opts, "oracle_noise_exit": ne, "verdict_mismatch": int((acc != oacc).sum()),
                          "max_rel": float(rel.max()), "n_above_1e-4": int((rel > 1e-4).sum()),
                          "worst": [dict(L=int(info["window_len"][i]), gpu=float(info["max_chi2"][i]), orc=float(orep["max_chi2"][i]), it_gpu=int(info["iterations"][i]),
                                         it_orc=int(orep["iterations"][i]) if "iterations" in orep.dtype.names else -1, acc=int(acc[i])) for i in w]}))
    ipc.close()
