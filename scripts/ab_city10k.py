"""A/B on the City10000-shaped matrix (BASELINE cfg4, batch part): where should the one-warp streamed kernel hand over to the
256-thread kernel? Seeded sample of the 64.6 M checks (59 % of the windows are longer than 5400 edges)."""
import sys, os, time, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ipc_b200 import api, synth, sharding
n_s = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
config = sys.argv[2] if len(sys.argv) > 2 else "city10k"
caps = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
g, cfg = synth.make_config(config)
mem, cnd = api.pair_checks(g)
sel = np.sort(np.random.default_rng(0).choice(len(cnd), n_s, replace=False))
mem, cnd = mem[sel], cnd[sel]
L = sharding.window_lengths(g, mem, cnd)
dev = torch.device("cuda", 0)
md, cd = torch.from_numpy(mem).to(dev), torch.from_numpy(cnd).to(dev)
bits = torch.zeros((n_s + 31) // 32, dtype=torch.int32, device=dev)
variants = {
    "default (one warp per check up to 5400 edges, 256 threads beyond)": {},
    "one warp per check up to 7500": {"bucket4_cap": 7500},
    "one warp per check for every window (cap 10000)": {"bucket4_cap": 10000},
    "default caps, 512 threads per check beyond 5400": {"bucket5_nt": 512},
}
if caps:       # second form: python scripts/ab_city10k.py <checks> <config> <cap,cap,...>: hand-over length sweep
    variants = {f"one warp per check up to {c}": {"bucket4_cap": c} for c in caps}
print(json.dumps({"config": config, "checks": n_s, "windows_gt_5400": float((L > 5400).mean()), "L_median": float(np.median(L)), "L_pct_10_90": [float(x) for x in np.percentile(L, [10, 90])]}), flush=True)
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
    print(json.dumps({"variant": name, "opts": opts, "checks_per_s": n_s / min(ts[1:]), "times": [round(x, 3) for x in ts],
                      "bits_equal_first": bool((b == ref).all().item()), "bits_sha": hashlib.sha256(b.cpu().numpy().tobytes()).hexdigest()[:12]}), flush=True)
    ipc.close()
