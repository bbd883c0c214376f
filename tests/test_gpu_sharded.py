"""Multi-GPU path behind the C ABI (ipc_comm_init / ipc_check_batch_sharded / ipc_consistency_matrix_sharded), on the GPU box.
One GPU: a world-size-1 communicator exercises the NCCL binding (dlopen, ncclCommInitRank, in-place all-gather).
Two or more GPUs: one handle per device driven from threads of this process, results compared with the single-handle batch."""
import threading

import numpy as np
import pytest

from ipc_b200 import api, sharding, synth

pytestmark = pytest.mark.gpu


def _checks(scale=0.2, n=4001, seed=3):
    g, cfg = synth.make_config("intel", scale=scale)
    mem, cnd = api.pair_checks(g)
    if len(cnd) > n:
        sel = np.sort(np.random.default_rng(seed).choice(len(cnd), n, replace=False))
        mem, cnd = mem[sel], cnd[sel]
    return g, cfg, mem, cnd


def test_world1_communicator_and_sharded_batch(gpu_lib):
    g, cfg, mem, cnd = _checks()
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    want, _ = ipc.check_batch(mem, cnd, want_info=False)
    ipc.comm_init(rank=0, world=1, uid=api.comm_unique_id())
    assert ipc.comm_info()[:2] == (0, 1)
    parts, wpr = sharding.shard_plan(sharding.window_lengths(g, mem, cnd), 1)
    words = ipc.check_batch_sharded(mem, cnd, wpr + 3)                 # padding words beyond the shard must come back zero
    assert np.array_equal(sharding.decode_gathered(words[:, :wpr], parts, len(cnd)), want)
    assert not words[:, (len(cnd) + 31) // 32:].any()
    rows_a, order_a, solved_a = ipc.consistency_matrix()
    rows_b, order_b, solved_b = ipc.consistency_matrix(sharded=True)
    assert solved_a == solved_b and np.array_equal(order_a, order_b) and np.array_equal(rows_a, rows_b)
    ipc.close()


def test_two_gpu_sharded_batch_and_matrix(gpu_lib):
    world = min(gpu_lib.lib().ipc_device_count(), 4)
    if world < 2:
        pytest.skip("needs two or more GPUs")
    g, cfg, mem, cnd = _checks(n=4003)
    if len(cnd) % world == 0:                                           # shards of unequal size
        mem, cnd = mem[:-1], cnd[:-1]
    single = gpu_lib.IPC.from_graph(g, cfg)
    want, _ = single.check_batch(mem, cnd, want_info=False)
    rows_want, order_want, solved_want = single.consistency_matrix()
    single.close()
    parts, wpr = sharding.shard_plan(sharding.window_lengths(g, mem, cnd), world)
    assert len({len(p) for p in parts}) > 1
    uid = api.comm_unique_id()
    out, err = [None] * world, [None] * world

    def run(r):
        try:
            ipc = api.IPC.from_graph(g, cfg, device=r)
            ipc.comm_init(rank=r, world=world, uid=uid)
            words = ipc.check_batch_sharded(mem[parts[r]], cnd[parts[r]], wpr)
            rows, order, solved = ipc.consistency_matrix(sharded=True)
            out[r] = (words, rows, order, solved, ipc.comm_info())
            ipc.close()
        except Exception as e:      # noqa: BLE001 — surfaced below
            err[r] = e

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=600) for t in th]
    assert all(e is None for e in err), err
    for r in range(world):
        words, rows, order, solved, info = out[r]
        assert info[0] == r and info[1] == world and info[2] == 2      # exactly one all-gather per batch / matrix
        assert np.array_equal(sharding.decode_gathered(words, parts, len(cnd)), want)
        assert solved == solved_want and np.array_equal(order, order_want) and np.array_equal(rows, rows_want)


def _cli_case(tmp_path, name, scale):
    import os
    from ipc_b200 import g2o
    g, cfg = synth.make_config(name, scale=scale)
    ds, gt, out, yml = (str(tmp_path / f) for f in ("graph.g2o", "gt.txt", "res.txt", "cfg.yaml"))
    g2o.write_g2o(g, ds)
    g2o.write_trajectory(g.gt, gt)
    g2o.write_config(yml, name, ds, gt, out, g.n_true, cfg)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return g, cfg, yml, out, os.path.join(root, "cli")


def _matrix_cli(exe, yml, *extra):
    import re
    import subprocess
    r = subprocess.run([exe, "-c", yml, "--matrix", *extra], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    m = re.search(r"TP (\d+) FP (\d+) TN (\d+) FN (\d+)", r.stdout)
    s = re.search(r"(\d+) candidates, (\d+) solved checks on (\d+) GPU", r.stdout)
    return tuple(int(x) for x in m.groups()), tuple(int(x) for x in s.groups())


@pytest.mark.parametrize("name,scale,tester", [("intel", 0.3, "ipc_tester_2D"), ("sphere", 0.05, "ipc_tester_3D")])
def test_cli_matrix_mode_equals_api(gpu_lib, tmp_path, name, scale, tester):
    """`ipc_tester_* --matrix` (C++ host over the C ABI): consistency matrix + greedy consensus, same confusion counts as the Python mirror;
    with two or more GPUs `--gpus N` (one handle per device on host threads, NCCL all-gather behind the C ABI) gives the same answer."""
    import os
    import subprocess
    g, cfg, yml, out, cli = _cli_case(tmp_path, name, scale)
    subprocess.check_call(["make", "-s", "-C", cli])
    ipc = gpu_lib.IPC.from_graph(g, cfg)
    rows, order, solved = ipc.consistency_matrix()
    sel = ipc.greedy_consensus(rows)
    ipc.close()
    acc = np.zeros(g.n_loops, bool); acc[order[sel]] = True
    truth = np.arange(g.n_loops) < g.n_true
    want = (int((acc & truth).sum()), int((acc & ~truth).sum()), int((~acc & ~truth).sum()), int((~acc & truth).sum()))
    got, (n_c, n_solved, n_gpu) = _matrix_cli(os.path.join(cli, tester), yml)
    assert got == want and n_c == g.n_loops and n_solved == solved and n_gpu == 1
    pr = open(out[:-3] + "PR").read().split()
    assert abs(float(pr[0]) - want[0] / max(want[0] + want[1], 1)) < 1e-5
    world = min(gpu_lib.lib().ipc_device_count(), 4)
    if world >= 2:
        got_n, (_, n_solved_n, n_gpu_n) = _matrix_cli(os.path.join(cli, tester), yml, "--gpus", str(world))
        assert got_n == want and n_solved_n == solved and n_gpu_n == world
