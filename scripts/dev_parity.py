"""Dev script: GPU batch vs CPU oracle on a scaled config (run on the GPU box)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipc_b200 import synth, api
from oracle import pyoracle as po

name = sys.argv[1] if len(sys.argv) > 1 else "intel"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
max_checks = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
noise_exit = int(sys.argv[4]) if len(sys.argv) > 4 else 1
g, cfg = synth.make_config(name, scale=scale)
print(name, "poses", g.n_poses, "loops", g.n_loops, "true", g.n_true)
mem, cnd = api.pair_checks(g)
print("checks", len(cnd))
if len(cnd) > max_checks:
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(len(cnd), max_checks, replace=False))
    mem, cnd = mem[sel], cnd[sel]
ipc = api.IPC.from_graph(g, cfg)
ipc.set_option("noise_exit", noise_exit)
t = time.time(); acc, info = ipc.check_batch(mem, cnd); t_gpu = time.time() - t
t = time.time(); acc, info = ipc.check_batch(mem, cnd); t_gpu2 = time.time() - t
print("gpu first %.3fs second %.3fs -> %.0f checks/s" % (t_gpu, t_gpu2, len(cnd) / t_gpu2), "stats", ipc.last_batch_stats())
orc = po.OracleIPC(g, cfg)
ptr = np.zeros(len(cnd) + 1, dtype=np.int32); idx = []
for i, (m, c) in enumerate(zip(mem, cnd)):
    if m >= 0: idx.append(m)
    idx.append(c); ptr[i + 1] = len(idx)
t = time.time(); oacc, orep = orc.check_batch(ptr, np.array(idx, dtype=np.int32), n_threads=os.cpu_count()); t_cpu = time.time() - t
print("oracle %.2fs on %d threads -> %.1f checks/s" % (t_cpu, os.cpu_count(), len(cnd) / t_cpu))
mism = np.nonzero(acc != oacc)[0]
rel = np.abs(info["max_chi2"] - orep["max_chi2"]) / np.maximum(np.abs(orep["max_chi2"]), 1e-12)
relc = np.abs(info["cand_chi2"] - orep["cand_chi2"]) / np.maximum(np.abs(orep["cand_chi2"]), 1e-12)
print("verdict mismatches:", len(mism), "of", len(cnd), "| accepted gpu", acc.sum(), "oracle", oacc.sum())
print("max_chi2 rel err: median %.3g  p99 %.3g  max %.3g ; > 1e-4: %d" % (np.median(rel), np.quantile(rel, 0.99), rel.max(), (rel > 1e-4).sum()))
print("cand_chi2 rel err: max %.3g" % relc.max())
print("iterations gpu median %d max %d | oracle median %d max %d" % (np.median(info["iterations"]), info["iterations"].max(), np.median(orep["iterations"]), orep["iterations"].max()))
print("evals gpu median %d | oracle median %d" % (np.median(info["evals"]), np.median(orep["evals"])))
bad = np.argsort(-rel)[:8]
for i in bad:
    print("  chk", i, "m", mem[i], "c", cnd[i], "L", info["window_len"][i], "K", info["n_loops"][i], "gpu", info["max_chi2"][i], info["iterations"][i], info["evals"][i],
          "orc", orep["max_chi2"][i], orep["iterations"][i], orep["evals"][i], orep["result"][i], "acc", acc[i], oacc[i])
