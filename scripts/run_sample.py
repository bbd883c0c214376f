"""One batch of a seeded sample of a config's check list through the device-resident entry point (for ncu captures / A-B runs).
Usage: python scripts/run_sample.py <n_checks> <config> [opt=value ...]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ipc_b200 import api, synth
n_s = int(sys.argv[1]); name = sys.argv[2]
g, cfg = synth.make_config(name)
mem, cnd = api.pair_checks(g)
if n_s < len(cnd):
    sel = np.sort(np.random.default_rng(0).choice(len(cnd), n_s, replace=False))
    mem, cnd = mem[sel], cnd[sel]
n_s = len(cnd)
dev = torch.device("cuda", 0)
md, cd = torch.from_numpy(mem).to(dev), torch.from_numpy(cnd).to(dev)
bits = torch.zeros((n_s + 31) // 32, dtype=torch.int32, device=dev)
ipc = api.IPC.from_graph(g, cfg)
for kv in sys.argv[3:]:
    k, v = kv.split("="); ipc.set_option(k, float(v))
st = torch.cuda.current_stream()
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    ipc.check_batch_dev(n_s, md.data_ptr(), cd.data_ptr(), bits.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print(json.dumps({"checks": n_s, "checks_per_s": n_s / dt, "opts": sys.argv[3:]}))
