import json, os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ipc_b200 import api, synth
g, cfg = synth.make_config("intel")
order = g.time_order()
def clocks():
    try: return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,pstate,power.draw", "--format=csv,noheader"], capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception as e: return str(e)
for rep in range(2):
    ipc = api.IPC.from_graph(g, cfg, candidates=False)
    ts = []; ev = []; K = []
    t_all = time.perf_counter()
    for k, l in enumerate(order):
        t = time.perf_counter()
        ok, ci = ipc.agreementCheck((g.loop_from[l], g.loop_to[l], g.loop_meas[l], g.loop_info[l]))
        ts.append(time.perf_counter() - t); ev.append(ci.evals); K.append(ci.n_loops)
        if k in (5, 100, 300): print("  check", k, "clocks:", clocks(), flush=True)
    ts = np.array(ts); ev = np.array(ev); K = np.array(K)
    print(json.dumps({"rep": rep, "total_s": time.perf_counter() - t_all, "ms_per_eval_first50": 1e3 * ts[:50].sum() / max(1, ev[:50].sum()), "ms_per_eval_last50": 1e3 * ts[-50:].sum() / max(1, ev[-50:].sum()),
                      "K_last": int(K[-1]), "max_check_s": float(ts.max()), "argmax": int(ts.argmax()), "evals_at_max": int(ev[ts.argmax()])}), flush=True)
    ipc.close()
