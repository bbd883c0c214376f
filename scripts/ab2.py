import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ipc_b200 import api, synth
n_s = 200000
g, cfg = synth.make_config("m3500")
mem, cnd = api.pair_checks(g)
sel = np.sort(np.random.default_rng(0).choice(len(cnd), n_s, replace=False))
mem, cnd = mem[sel], cnd[sel]
dev = torch.device("cuda", 0)
md, cd = torch.from_numpy(mem).to(dev), torch.from_numpy(cnd).to(dev)
bits = torch.zeros((n_s + 31) // 32, dtype=torch.int32, device=dev)
ipc = api.IPC.from_graph(g, cfg)
st = torch.cuda.current_stream()
ts = []
for rep in range(4):
    torch.cuda.synchronize(); t = time.perf_counter()
    ipc.check_batch_dev(n_s, md.data_ptr(), cd.data_ptr(), bits.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
print(json.dumps({"lib": api.LIB_PATH, "checks_per_s": n_s / min(ts[1:]), "times": [round(x, 3) for x in ts], "popcount": int(sum(bin(int(x) & 0xffffffff).count("1") for x in bits.cpu().numpy()[:2000]))}))
