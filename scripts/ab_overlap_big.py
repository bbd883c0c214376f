"""Same A/B as ab_overlap.py on ONE large seeded sample of the M3500 matrix, alternating, one run each (steady-state check)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ipc_b200 import api, synth
n_s = int(sys.argv[1]) if len(sys.argv) > 1 else 600000
dev = torch.device("cuda", 0)
g, cfg = synth.make_config("m3500")
mem, cnd = api.pair_checks(g)
sel = np.sort(np.random.default_rng(0).choice(len(cnd), n_s, replace=False))
md, cd = torch.from_numpy(mem[sel]).to(dev), torch.from_numpy(cnd[sel]).to(dev)
bits = torch.zeros((n_s + 31) // 32, dtype=torch.int32, device=dev)
ipc = api.IPC.from_graph(g, cfg)
st = torch.cuda.current_stream()
for ov in (1, 1, 0, 1, 0):
    ipc.set_option("overlap_buckets", ov)
    torch.cuda.synchronize(); t = time.perf_counter()
    ipc.check_batch_dev(n_s, md.data_ptr(), cd.data_ptr(), bits.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(json.dumps({"checks": n_s, "overlap_buckets": ov, "checks_per_s": n_s / dt, "s": round(dt, 4), "kernel_ms": ipc.last_kernel_ms()}), flush=True)
ipc.close()
