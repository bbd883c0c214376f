"""Where the SE(2) check kernel spends its cycles: runs a seeded sample of the M3500 check list through the profiling build
(make -C ipc_b200/csrc prof -> libipc_b200_prof.so, per-phase clock64 counters of every CTA's thread 0) and prints the split.
Usage: IPC_B200_LIB=ipc_b200/libipc_b200_prof.so python scripts/phase_clocks.py [n_checks]"""
import sys, os, json, ctypes as C
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
lib = C.CDLL(api.LIB_PATH)
names = ["setup+dead-reckon", "sweep: edge loop", "sweep: scan+barrier", "loop edges + GN solve (warp 0) + barrier", "norm pass", "gradient pass(es)",
         "rollback", "decision logic / other"]
ipc = api.IPC.from_graph(g, cfg)
st = torch.cuda.current_stream()
for rep in range(2):
    buf = (C.c_ulonglong * 16)()
    lib.ipc_debug_phase_clocks(buf, 1)
    torch.cuda.synchronize()
    ipc.check_batch_dev(n_s, md.data_ptr(), cd.data_ptr(), bits.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize()
lib.ipc_debug_phase_clocks(buf, 1)
v = [int(x) for x in buf]
tot = sum(v[:8])
out = {"kernel_ms": ipc.last_kernel_ms(), "checks": v[13], "mean_window": v[14] / max(1, v[13]), "evals_per_check": v[15] / max(1, v[13]),
       "per_check": {"gn_sweeps": v[8] / v[13], "blend_sweeps": v[9] / v[13], "relin_sweeps": v[10] / v[13], "norm_passes": v[11] / v[13], "gradient_passes": v[12] / v[13]},
       "cycle_share": {names[i]: round(v[i] / tot, 4) for i in range(8)},
       "kcycles_per_call": {"sweep loop": v[1] / max(1, v[8] + v[9] + v[10]) / 1e3, "sweep scan": v[2] / max(1, v[8] + v[9] + v[10]) / 1e3,
                            "solve": v[3] / max(1, v[15] + v[13]) / 1e3, "norm": v[4] / max(1, v[11]) / 1e3, "gradient": v[5] / max(1, v[12]) / 1e3}}
print(json.dumps(out, indent=1))
