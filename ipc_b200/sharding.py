"""Sharding of independent checks across the GPUs of one box.

Fast and pair checks are independent units (each starts from the dead-reckoned state), so the check list is dealt to the
ranks by estimated cost (window length, longest first, round robin) with NO data-path collective; the only exchange step is
ONE all_gather of the packed verdict words (N_c^2 / 8 bytes in total for a whole consistency matrix: latency bound), after
which every rank holds every verdict. One process per GPU, `torch.distributed` (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def window_lengths(graph, member, cand):
    """L = hi - lo of every check (union window when the member overlaps, src/consensus.cpp:157-159)."""
    a = np.minimum(graph.loop_from, graph.loop_to)
    b = np.maximum(graph.loop_from, graph.loop_to)
    member, cand = np.asarray(member), np.asarray(cand)
    m = np.where(member >= 0, member, cand)
    ov = (member >= 0) & ((np.minimum(b[m], b[cand]) - np.maximum(a[m], a[cand])) > 0)
    lo = np.where(ov, np.minimum(a[m], a[cand]), a[cand])
    hi = np.where(ov, np.maximum(b[m], b[cand]), b[cand])
    return (hi - lo).astype(np.int64)


def partition(cost, world: int, rank: int) -> np.ndarray:
    """Indices of the checks owned by `rank`: sort by cost (descending, stable), deal round robin. Deterministic, and the
    union over ranks is a permutation of range(len(cost))."""
    order = np.argsort(-np.asarray(cost), kind="stable")
    return np.sort(order[rank::world])


def pack_bits(v: np.ndarray) -> np.ndarray:
    n = len(v)
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(v)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return words


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    i = np.arange(n)
    return ((words[i >> 5] >> (i & 31).astype(np.uint32)) & 1).astype(bool)


def shard_plan(cost, world: int):
    """(parts, words_per_rank): parts[r] = check indices of rank r (partition), words_per_rank = 32-bit words that hold the
    largest shard — EVERY rank pads its verdict words to this count so the all-gather has equal contributions."""
    parts = [partition(cost, world, r) for r in range(world)]
    per = max((len(p) for p in parts), default=0)
    return parts, max(1, (per + 31) // 32)


def decode_gathered(gathered, parts, n: int) -> np.ndarray:
    """Verdict of every check of the whole list from the gathered words [world][words_per_rank]."""
    gathered = np.asarray(gathered).view(np.uint32).reshape(len(parts), -1)
    out = np.zeros(n, dtype=bool)
    for r, idx in enumerate(parts):
        out[idx] = unpack_bits(gathered[r], len(idx))
    return out


def sharded_verdicts(cost, compute_local, world: int, rank: int, dist=None, device=None):
    """Run `compute_local(indices) -> bool verdicts` on this rank's shard and all_gather the packed words.
    Returns the verdict of every check (same array on every rank)."""
    n = len(cost)
    mine = partition(cost, world, rank)
    local = np.asarray(compute_local(mine), dtype=bool)
    if world == 1 or dist is None:
        out = np.zeros(n, dtype=bool)
        out[mine] = local
        return out
    import torch
    parts, words = shard_plan(cost, world)               # shards differ in size by at most one check: pad to the largest
    buf = torch.zeros(words, dtype=torch.int32, device=device)
    buf[: (len(local) + 31) // 32] = torch.from_numpy(pack_bits(local).view(np.int32)).to(device)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)                       # the single collective of the path
    return decode_gathered(np.stack([t.cpu().numpy() for t in gathered]), parts, n)


def matrix_from_verdicts(graph, member, cand, verdict, order=None):
    """Dense boolean consistency matrix in time order from the verdicts of the solved checks (diagonal = fast check,
    overlapping pair = its K = 2 check, anything else = AND of the two diagonals)."""
    order = graph.time_order() if order is None else np.asarray(order)
    n = len(order)
    pos = np.empty(graph.n_loops, dtype=np.int64)
    pos[order] = np.arange(n)
    member, cand, verdict = np.asarray(member), np.asarray(cand), np.asarray(verdict, dtype=bool)
    diag = np.zeros(n, dtype=bool)
    d = member < 0
    diag[pos[cand[d]]] = verdict[d]
    M = np.logical_and.outer(diag, diag)
    p = ~d
    M[pos[member[p]], pos[cand[p]]] = verdict[p]
    M[pos[cand[p]], pos[member[p]]] = verdict[p]
    M[np.arange(n), np.arange(n)] = diag
    return M


def consistency_matrix_sharded(ipc, graph, world: int = 1, rank: int = 0, dist=None, device=None):
    """Pairwise consistency matrix with the solved checks dealt across `world` ranks (one process per GPU, each holding the
    whole graph and its own `ipc` handle): local ipc.check_batch on this rank's shard, ONE all_gather of the packed verdict
    words, then every rank assembles the same dense boolean matrix (time order). Returns (matrix, order)."""
    from . import api
    order = graph.time_order()
    member, cand = api.pair_checks(graph, order)
    cost = window_lengths(graph, member, cand)

    def local(idx):
        acc, _ = ipc.check_batch(member[idx], cand[idx], want_info=False)
        return acc

    verdict = sharded_verdicts(cost, local, world, rank, dist=dist, device=device)
    return matrix_from_verdicts(graph, member, cand, verdict, order), order
