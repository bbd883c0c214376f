"""N > 1 path on CPU: world_size 2 over gloo. The per-rank compute is replaced by a deterministic stand-in (no GPU here);
what is tested is the partition, the single all_gather of packed verdict words and the reassembly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ipc_b200 import api, sharding, synth


def _fake_verdict(member, cand):
    return ((member.astype(np.int64) * 7 + cand.astype(np.int64) * 13) % 5) < 2


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, _ = synth.make_config("intel", scale=0.25)
    mem, cnd = api.pair_checks(g)
    cost = sharding.window_lengths(g, mem, cnd)
    v = sharding.sharded_verdicts(cost, lambda idx: _fake_verdict(mem[idx], cnd[idx]), world, rank, dist=dist, device="cpu")
    np.save(os.path.join(out_dir, f"v{rank}.npy"), v)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g, _ = synth.make_config("intel", scale=0.25)
    mem, cnd = api.pair_checks(g)
    want = _fake_verdict(mem, cnd)
    v0, v1 = np.load(tmp_path / "v0.npy"), np.load(tmp_path / "v1.npy")
    assert np.array_equal(v0, want) and np.array_equal(v1, want)


def _worker_unequal(rank, world, port, out_dir, n):
    """bench.py's gather plan with UNEQUAL per-rank check counts (n not a multiple of world): every rank pads its words to
    shard_plan's words_per_rank, so the all_gather contributions have equal size (the round-1 bench hung exactly here)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    cost = rng.integers(2, 3000, n)
    truth = rng.random(n) < 0.5
    parts, wpr = sharding.shard_plan(cost, world)
    mine = parts[rank]
    buf = torch.zeros(wpr, dtype=torch.int32)
    w = sharding.pack_bits(truth[mine]).view(np.int32)
    buf[: len(w)] = torch.from_numpy(w)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    got = sharding.decode_gathered(np.stack([t.numpy() for t in gathered]), parts, n)
    np.save(os.path.join(out_dir, f"u{rank}.npy"), np.concatenate([[len(mine)], got.astype(np.int64)]))
    dist.barrier()
    dist.destroy_process_group()


def test_unequal_shards_gather(tmp_path):
    n = 1001                                  # 2 ranks: 501 + 500 checks
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker_unequal, args=(2, port, str(tmp_path), n), nprocs=2, join=True)
    truth = np.random.default_rng(5); truth.integers(2, 3000, n); want = truth.random(n) < 0.5
    u0, u1 = np.load(tmp_path / "u0.npy"), np.load(tmp_path / "u1.npy")
    assert u0[0] != u1[0] and u0[0] + u1[0] == n
    assert np.array_equal(u0[1:].astype(bool), want) and np.array_equal(u1[1:].astype(bool), want)


def test_shard_plan_pads_to_the_largest_shard():
    for n, world in [(1001, 2), (1000, 3), (7, 8), (2242424, 8)]:
        parts, wpr = sharding.shard_plan(np.arange(n) % 977, world)
        assert sum(len(p) for p in parts) == n and len(parts) == world
        assert wpr * 32 >= max(len(p) for p in parts) and wpr >= 1
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_partition_is_a_balanced_permutation():
    g, _ = synth.make_config("m3500", scale=0.2)
    mem, cnd = api.pair_checks(g)
    cost = sharding.window_lengths(g, mem, cnd)
    parts = [sharding.partition(cost, 8, r) for r in range(8)]
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(len(cost)))
    loads = np.array([cost[p].sum() for p in parts], dtype=np.float64)
    assert loads.max() / loads.min() < 1.01


def test_pack_unpack_roundtrip_and_matrix_assembly():
    rng = np.random.default_rng(0)
    v = rng.random(1000) < 0.4
    assert np.array_equal(sharding.unpack_bits(sharding.pack_bits(v), 1000), v)
    g, _ = synth.make_config("intel", scale=0.1)
    mem, cnd = api.pair_checks(g)
    ver = _fake_verdict(mem, cnd)
    M = sharding.matrix_from_verdicts(g, mem, cnd, ver)
    assert np.array_equal(M, M.T)
    o = g.time_order()
    pos = {int(l): k for k, l in enumerate(o)}
    for m, c, a in zip(mem, cnd, ver):
        if m >= 0:
            assert M[pos[int(m)], pos[int(c)]] == a
