import os

import numpy as np

from ipc_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    g = synth.Graph(int(z["dim"]), int(z["n_poses"]), z["odom_meas"], z["odom_info"], z["loop_from"], z["loop_to"], z["loop_meas"],
                    z["loop_info"], int(z["n_true"]))
    c = z["cfg"]
    cfg = dict(s_factor=float(c[0]), fast_reject_th=float(c[1]), slow_reject_th=float(c[2]), fast_reject_iter_base=int(c[3]),
               slow_reject_iter_base=int(c[4]))
    return z, g, cfg


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-9)
