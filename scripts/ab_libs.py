"""A/B of two builds of libipc_b200.so on the same box, same seeded sample of the M3500 check list (device-resident), one process per
library (IPC_B200_LIB selects it). Usage: python scripts/ab_libs.py <n_checks> <lib_a.so> <lib_b.so> [config] [opt=value ...]"""
import sys, os, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time, json
sys.path.insert(0, %r)
import numpy as np, torch
from ipc_b200 import api, synth
n_s = int(sys.argv[1]); name = sys.argv[2]
g, cfg = synth.make_config(name)
mem, cnd = api.pair_checks(g)
lo_hi = os.environ.get("AB_WINDOW")          # "lo,hi": keep the checks whose window length is in (lo, hi]
if lo_hi:
    a = np.minimum(g.loop_from, g.loop_to); b = np.maximum(g.loop_from, g.loop_to)
    ca, cb = a[cnd], b[cnd]
    ma, mb = np.where(mem >= 0, a[np.maximum(mem, 0)], ca), np.where(mem >= 0, b[np.maximum(mem, 0)], cb)
    ov = (np.minimum(cb, mb) - np.maximum(ca, ma)) > 0
    L = np.where(ov, np.maximum(cb, mb) - np.minimum(ca, ma), cb - ca)
    lo_, hi_ = (int(x) for x in lo_hi.split(","))
    keep = (L > lo_) & (L <= hi_)
    mem, cnd = mem[keep], cnd[keep]
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
ts = []
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    ipc.check_batch_dev(n_s, md.data_ptr(), cd.data_ptr(), bits.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
import hashlib
print(json.dumps({"lib": os.environ.get("IPC_B200_LIB"), "opts": sys.argv[3:], "checks": n_s, "checks_per_s": n_s / min(ts[1:]), "times": [round(x, 3) for x in ts],
                  "bits_sha": hashlib.sha1(bits.cpu().numpy().tobytes()).hexdigest()[:12]}), flush=True)
''' % ROOT
n = sys.argv[1]; libs = sys.argv[2:4]; name = sys.argv[4] if len(sys.argv) > 4 else "m3500"; opts = sys.argv[5:]
# option sets are separated by "--": every set is run on every library
sets, cur = [], []
for o in opts:
    if o == "--": sets.append(cur); cur = []
    else: cur.append(o)
sets.append(cur)
for rnd in range(2):
    for lib in dict.fromkeys(libs):
        for st in sets:
            env = dict(os.environ, IPC_B200_LIB=os.path.abspath(lib))
            subprocess.run([sys.executable, "-c", CHILD, n, name] + st, env=env, check=False)
