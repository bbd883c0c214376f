"""A/B: the bucket launches of one batch side by side on their own streams (default) against one after the other on the caller's
stream (option overlap_buckets = 0). Seeded samples of the M3500 and City10000-shaped matrices; verdict words must be identical."""
import sys, os, time, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ipc_b200 import api, synth
dev = torch.device("cuda", 0)
for config, n_s in (("m3500", 200000), ("m3500", 25000), ("city10k", 40000)):
    g, cfg = synth.make_config(config)
    mem, cnd = api.pair_checks(g)
    sel = np.sort(np.random.default_rng(0).choice(len(cnd), n_s, replace=False))
    md, cd = torch.from_numpy(mem[sel]).to(dev), torch.from_numpy(cnd[sel]).to(dev)
    bits = torch.zeros((n_s + 31) // 32, dtype=torch.int32, device=dev)
    ipc = api.IPC.from_graph(g, cfg)
    st = torch.cuda.current_stream()
    for ov in (1, 0, 1, 0):
        ipc.set_option("overlap_buckets", ov)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); t = time.perf_counter()
            ipc.check_batch_dev(n_s, md.data_ptr(), cd.data_ptr(), bits.data_ptr(), None, st.cuda_stream)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
        print(json.dumps({"config": config, "checks": n_s, "overlap_buckets": ov, "checks_per_s": n_s / min(ts[1:]), "times": [round(x, 4) for x in ts],
                          "kernel_ms": ipc.last_kernel_ms(), "bits_sha": hashlib.sha256(bits.cpu().numpy().tobytes()).hexdigest()[:12]}), flush=True)
    ipc.close()
