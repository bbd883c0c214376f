"""A/B timing of kernel options on a seeded sample of the M3500 workload (device-resident check list)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ipc_b200 import api, synth
n_s = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
g, cfg = synth.make_config("m3500")
mem, cnd = api.pair_checks(g)
sel = np.sort(np.random.default_rng(0).choice(len(cnd), n_s, replace=False))
mem, cnd = mem[sel], cnd[sel]
dev = torch.device("cuda", 0)
md, cd = torch.from_numpy(mem).to(dev), torch.from_numpy(cnd).to(dev)
bits = torch.zeros((n_s + 31) // 32, dtype=torch.int32, device=dev)
variants = {
    "default (sd_fuse=2; 32x16, 64x8, 128x3, 128x2, 256x1)": {},
    "sd_fuse=0 (norm pass, then gradient pass)": {"sd_fuse": 0},
    "sd_fuse=1 (always the gradient pass)": {"sd_fuse": 1},
    "b2_cap1700": {"bucket2_cap": 1700},
    "b2_cap1000": {"bucket2_cap": 1000},
    "b3_192x2": {"bucket3_nt": 192, "bucket3_minb": 2},
    "b3_256x1": {"bucket3_nt": 256, "bucket3_minb": 1},
    "b4_384x1": {"bucket4_nt": 384, "bucket4_minb": 1},
    "b1_64x4": {"bucket1_minb": 4},
}
ref = None
for name, opts in variants.items():
    ipc = api.IPC.from_graph(g, cfg)
    for k, v in opts.items(): ipc.set_option(k, v)
    st = torch.cuda.current_stream()
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        ipc.check_batch_dev(n_s, md.data_ptr(), cd.data_ptr(), bits.data_ptr(), None, st.cuda_stream)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    b = bits.clone()
    if ref is None: ref = b
    print(json.dumps({"variant": name, "checks_per_s": n_s / min(ts[1:]), "times": [round(x, 3) for x in ts], "bits_equal_first": bool((b == ref).all().item())}), flush=True)
    ipc.close()
