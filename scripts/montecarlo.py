#!/usr/bin/env python
"""Monte-Carlo evaluation protocol of the reference as one command.

Reproduces /root/reference/bash/ipc_experiments_2D.sh:3-41 (and _3D.sh) + scripts/generateDataset.py:188-246:

    for dataset, for outliers in 10 20 ... 100, for run in 00 .. 09:
        SPOILED_DATA/<outliers>/<run>.g2o  = clean graph + <outliers> random loop edges (generateDataset.py rule)
        <opt>_<outliers>_<run>.yaml        = dataset params with dataset/output replaced and
                                             s_factor 10 (2D) / 50 (3D), k_buddies 2 (2D), use_best_k_buddies false,
                                             use_recovery true, fast/slow_reject_th 10.64 (2D only)
        ipc_tester_2D|3D -c <yaml>         (the ten runs of one outlier level side by side, like the script's `&` ... `wait`)
    -> EXP/<date>/<opt>/<outliers>/<run>.TRJ + .PR ("<precision> <recall>\\n<total_s> <avg_s>\\n", src/simulation.cpp:101-104)

and collects every .PR into summary.json (mean / min precision and recall and checks/s per outlier level).
The clean graph is one of the synthetic named shapes (no dataset ships with the reference, no network) or any g2o file
(--g2o, with --canonic = number of true loops = the file's loop count).

  python scripts/montecarlo.py --dataset intel --workdir /tmp/mc                    # the 10 x 10 protocol, stream (ipc_tester_2D)
  python scripts/montecarlo.py --dataset sphere --scale 0.1 --outliers 10,20 --runs 3
  python scripts/montecarlo.py --dataset intel --mode matrix                        # same spoiled graphs through the pair batch
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ipc_b200 import g2o, spoil, synth  # noqa: E402

OVERRIDES_2D = dict(s_factor=10.0, k_buddies=2, use_best_k_buddies=False, use_recovery=True, fast_reject_th=10.64, slow_reject_th=10.64)
OVERRIDES_3D = dict(s_factor=50.0, use_best_k_buddies=False, use_recovery=True)      # _3D.sh sets neither k_buddies nor the thresholds


def run_yaml(path, name, dataset, gt, output, canonic, cfg, dim):
    """params.yaml + the `yq -i` edits of bash/ipc_experiments_{2D,3D}.sh:24-33."""
    ov = OVERRIDES_2D if dim == 2 else OVERRIDES_3D
    c = dict(cfg)
    c.update({k: v for k, v in ov.items() if k in ("s_factor", "fast_reject_th", "slow_reject_th")})
    b = lambda x: "true" if x else "false"  # noqa: E731
    lines = [f'name: "{name}"', f'dataset: "{dataset}"', f'ground_truth: "{gt}"', f'output: "{output}"', "visualize: false", f"canonic_inliers: {canonic}",
             f"fast_reject_th: {c['fast_reject_th']}", f"fast_reject_iter_base: {c['fast_reject_iter_base']}", f"slow_reject_th: {c['slow_reject_th']}",
             f"slow_reject_iter_base: {c['slow_reject_iter_base']}", f"s_factor: {c['s_factor']}", f"k_buddies: {ov.get('k_buddies', 2)}",
             f"use_best_k_buddies: {b(ov['use_best_k_buddies'])}", f"use_recovery: {b(ov['use_recovery'])}"]
    open(path, "w").write("\n".join(lines) + "\n")
    return c


def read_pr(path):
    t = open(path).read().split()
    return dict(precision=float(t[0]), recall=float(t[1]), total_s=float(t[2]), avg_s=float(t[3]))


def matrix_run(g, cfg, device):
    """--mode matrix: the spoiled graph through the pair batch (consistency matrix + greedy consensus) instead of the stream."""
    from ipc_b200 import api
    ipc = api.IPC.from_graph(g, cfg, device=device)
    t0 = time.perf_counter()
    rows, order, solved = ipc.consistency_matrix()             # rows in time order: order[k] = file index of row k
    sel = ipc.greedy_consensus(rows)
    dt = time.perf_counter() - t0
    ipc.close()
    truth = np.arange(g.n_loops) < g.n_true
    acc = np.zeros(g.n_loops, bool)
    acc[np.asarray(order)[np.asarray(sel, dtype=bool)]] = True
    tp, fp, fn = int((acc & truth).sum()), int((acc & ~truth).sum()), int((~acc & truth).sum())
    return dict(precision=tp / max(tp + fp, 1), recall=tp / max(tp + fn, 1), total_s=dt, avg_s=dt / max(int(solved), 1), checks=int(solved))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dataset", default="intel", help="named synthetic shape (intel, m3500, sphere, city10k, synth50k) — ignored with --g2o")
    ap.add_argument("--g2o", help="clean g2o file to spoil instead of a synthetic shape")
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--gt", help="ground-truth trajectory file for --g2o (one pose per line); a zero trajectory is written if absent")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--outliers", default="10,20,30,40,50,60,70,80,90,100")
    ap.add_argument("--runs", type=int, default=10)
    ap.add_argument("--jobs", type=int, default=10, help="testers side by side per outlier level (the script starts all ten)")
    ap.add_argument("--workdir", default="montecarlo_out")
    ap.add_argument("--opt", default="B200_IPC")
    ap.add_argument("--date", default=time.strftime("%d%m%y"))
    ap.add_argument("--mode", choices=["stream", "matrix"], default="stream")
    ap.add_argument("--spoiler", choices=["file", "numpy"], default="file",
                    help="file: ipc_b200.spoil — the clean g2o is spoiled on disk exactly like generateDataset.py -n <outliers> --seed <s> "
                         "(byte-identical output, pinned by tests/golden/spoil); numpy: synth.add_outliers on the in-memory graph")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--tester", help="tester executable (default cli/ipc_tester_2D|3D)")
    ap.add_argument("--tester-args", default="--quiet", help="extra arguments for the tester")
    a = ap.parse_args(argv)

    if a.g2o:
        dim = a.dim or 2
        clean = g2o.read_g2o(a.g2o, dim)
        name = os.path.splitext(os.path.basename(a.g2o))[0]
        base_cfg = dict(synth.CONFIGS["intel" if dim == 2 else "sphere"]["cfg"])
    else:
        clean = synth.make_clean(a.dataset, a.scale)
        dim, name, base_cfg = clean.dim, a.dataset.upper(), dict(synth.CONFIGS[a.dataset]["cfg"])
    tester = a.tester or os.path.join(ROOT, "cli", "ipc_tester_2D" if dim == 2 else "ipc_tester_3D")
    root = os.path.join(os.path.abspath(a.workdir), name)
    gt_path = os.path.join(root, "gt.txt")
    os.makedirs(root, exist_ok=True)
    if a.gt:
        gt_path = os.path.abspath(a.gt)
    else:
        g2o.write_trajectory(clean.gt if clean.gt is not None else np.zeros((clean.n_poses, 3 if dim == 2 else 7)), gt_path)
    clean_path = os.path.abspath(a.g2o) if a.g2o else os.path.join(root, "clean.g2o")
    if not a.g2o:
        g2o.write_g2o(clean, clean_path)
    levels = [int(x) for x in a.outliers.split(",") if x]
    summary = dict(dataset=name, dim=dim, n_poses=clean.n_poses, true_loops=clean.n_loops, mode=a.mode, runs=a.runs, levels={})
    t_all = time.perf_counter()
    for out in levels:
        spoiled_dir = os.path.join(root, "SPOILED_DATA", str(out))
        exp_dir = os.path.join(root, "EXP", a.date, a.opt, str(out))
        os.makedirs(spoiled_dir, exist_ok=True); os.makedirs(exp_dir, exist_ok=True)
        jobs, results = [], []
        for run in range(a.runs):
            tag = f"{run:02d}"
            ds = os.path.join(spoiled_dir, tag + ".g2o")
            seed = 100000 + 1000 * out + run
            if a.spoiler == "file":
                spoil.spoil_g2o(clean_path, ds, outliers=out, seed=seed)
                g = g2o.read_g2o(ds, dim, n_true=clean.n_loops) if a.mode == "matrix" else None
            else:
                g = synth.add_outliers(clean, out, seed=seed)
                g2o.write_g2o(g, ds)
            trj = os.path.join(exp_dir, tag + ".TRJ")
            yml = os.path.join(root, f"{a.opt}_{out}_{tag}.yaml")
            cfg = run_yaml(yml, name, ds, gt_path, trj, clean.n_loops, base_cfg, dim)
            if a.mode == "matrix":
                results.append(matrix_run(g, cfg, a.device))
                os.remove(yml)
                continue
            jobs.append((tag, yml, trj, subprocess.Popen([tester, "-c", yml, "--device", str(a.device)] + a.tester_args.split(),
                                                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
            if len(jobs) >= a.jobs or run + 1 == a.runs:            # `wait`
                for tag_, yml_, trj_, p in jobs:
                    so, se = p.communicate()
                    if p.returncode != 0:
                        raise SystemExit(f"tester failed on {yml_}: {se.strip()}")
                    results.append(read_pr(trj_[:-3] + "PR"))
                    os.remove(yml_)                                     # `rm ./*.yaml`
                jobs = []
        pr = np.array([[r["precision"], r["recall"], r["total_s"]] for r in results])
        n_cand = clean.n_loops + out
        summary["levels"][str(out)] = dict(precision_mean=float(pr[:, 0].mean()), precision_min=float(pr[:, 0].min()), recall_mean=float(pr[:, 1].mean()),
                                           recall_min=float(pr[:, 1].min()), total_s_mean=float(pr[:, 2].mean()),
                                           candidates=n_cand, candidates_per_s=float(n_cand / max(pr[:, 2].mean(), 1e-12)), runs=len(results))
        print(f"[{name}] outliers {out:4d}: precision {pr[:, 0].mean():.4f} (min {pr[:, 0].min():.4f})  recall {pr[:, 1].mean():.4f}  "
              f"{pr[:, 2].mean():.3f} s per run", flush=True)
    summary["wall_s"] = time.perf_counter() - t_all
    with open(os.path.join(root, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print("Finished " + name)
    return summary


if __name__ == "__main__":
    main()
