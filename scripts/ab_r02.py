"""Round 2 A/B: odometry window staged in shared memory by cp.async.bulk + mbarrier (MODE 2 kernels, option stage_odom) against the
default kernels, on a seeded sample of the M3500 workload (device-resident check list)."""
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
    "default (state in smem, odometry through the slot-ordered L2 scratch)": {},
    "stage_odom=1 (cp.async.bulk window staging; caps 1100x3 / 1700x2 / 3500x1)": {"stage_odom": 1},
    "stage_odom=1, 1100-1700 on 128x2 and the rest on 256x1 -> try 192-free split: b3 cap 1700, b4 cap 3500": {"stage_odom": 1, "bucket2_cap": 800},
    "default with the same caps as the staged table (1100 / 1700 / 3500)": {"bucket2_cap": 1100, "bucket3_cap": 1700, "bucket4_cap": 3500},
    "default, b2_cap1700": {"bucket2_cap": 1700},
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
