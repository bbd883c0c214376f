"""Host-side mirror of the reference's IPC interface over the C ABI (include/ipc_b200.h).

``IPC`` has the method names and argument meaning of the reference class template
``IPC<EDGE, VERTEX>`` (/root/reference/include/ipc/consensus.hpp:5-33): ``agreementCheck``,
``removeEdgeFromCnS``, ``addEdgeToCnS``, ``getMaxConsensusSet`` — an edge is the tuple
``(from, to, measurement, information)`` instead of a ``g2o::EdgeSE2*``. The batched entry points
(``check_batch``, ``consistency_matrix``) are the throughput path. Everything is computed by
``libipc_b200.so`` on the GPU; if the library or a CUDA device is missing the constructor raises —
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPC_B200_LIB", os.path.join(_HERE, "libipc_b200.so"))
_LIB = None

# every symbol include/ipc_b200.h declares
SYMBOLS = ["ipc_last_error", "ipc_device_count", "ipc_create", "ipc_destroy", "ipc_agreement_check", "ipc_remove_edge",
           "ipc_add_edge", "ipc_consensus_size", "ipc_get_consensus", "ipc_get_poses", "ipc_final_optimize", "ipc_set_candidates", "ipc_check_batch",
           "ipc_check_batch_dev", "ipc_last_batch_stats", "ipc_last_kernel_ms", "ipc_consistency_matrix", "ipc_greedy_consensus", "ipc_set_option",
           "ipc_comm_unique_id", "ipc_comm_init", "ipc_comm_info", "ipc_check_batch_sharded", "ipc_check_batch_sharded_dev",
           "ipc_consistency_matrix_sharded", "ipc_stream_profile", "ipc_agreement_check_stream"]


class IpcError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("s_factor", C.c_double), ("fast_reject_th", C.c_double), ("slow_reject_th", C.c_double),
                ("fast_reject_iter_base", C.c_int), ("slow_reject_iter_base", C.c_int)]


class CheckInfo(C.Structure):
    _fields_ = [("max_chi2", C.c_double), ("cand_chi2", C.c_double), ("sum_chi2", C.c_double), ("iterations", C.c_int),
                ("evals", C.c_int), ("window_len", C.c_int), ("n_loops", C.c_int)]


INFO_DTYPE = np.dtype([("max_chi2", "f8"), ("cand_chi2", "f8"), ("sum_chi2", "f8"), ("iterations", "i4"), ("evals", "i4"),
                       ("window_len", "i4"), ("n_loops", "i4")])
assert INFO_DTYPE.itemsize == C.sizeof(CheckInfo)


def lib():
    """Load libipc_b200.so (built by ``__graft_entry__.build()`` / ``make -C ipc_b200/csrc``). Fails loudly."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise IpcError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` — there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.ipc_last_error.restype = C.c_char_p
        L.ipc_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
        L.ipc_destroy.argtypes = [C.c_void_p]
        L.ipc_destroy.restype = None
        L.ipc_agreement_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(CheckInfo)]
        L.ipc_remove_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ipc_add_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ipc_consensus_size.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ipc_get_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ipc_get_poses.argtypes = [C.c_void_p, C.c_void_p]
        if hasattr(L, "ipc_final_optimize"):
            L.ipc_final_optimize.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ipc_set_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ipc_check_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ipc_check_batch_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ipc_last_batch_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.ipc_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.ipc_consistency_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.ipc_greedy_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ipc_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ipc_stream_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ipc_agreement_check_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ipc_comm_unique_id.argtypes = [C.c_void_p]
        L.ipc_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ipc_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        L.ipc_check_batch_sharded.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ipc_check_batch_sharded_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ipc_consistency_matrix_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        _LIB = L
    return _LIB


def _chk(rc):
    if rc != 0:
        raise IpcError(f"ipc_b200 error {rc}: {lib().ipc_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class IPC:
    """Drop-in for ``IPC<EDGE, VERTEX>`` (include/ipc/consensus.hpp:5-33) on flat arrays."""

    def __init__(self, dim: int, odom_meas, odom_info, cfg: dict, device: int = 0):
        self.dim = dim
        self.d = 3 if dim == 2 else 6
        self.mw = 3 if dim == 2 else 7
        om, oi = _f64(odom_meas), _f64(odom_info)
        self.n_poses = om.shape[0] + 1
        c = Config(cfg["s_factor"], cfg["fast_reject_th"], cfg["slow_reject_th"], cfg["fast_reject_iter_base"], cfg["slow_reject_iter_base"])
        h = C.c_void_p()
        _chk(lib().ipc_create(dim, self.n_poses, _p(om), _p(oi), C.byref(c), device, C.byref(h)))
        self._h = h
        self.n_candidates = 0

    @classmethod
    def from_graph(cls, graph, cfg: dict, device: int = 0, candidates: bool = True):
        o = cls(graph.dim, graph.odom_meas, graph.odom_info, cfg, device)
        if candidates:
            o.set_candidates(graph.loop_from, graph.loop_to, graph.loop_meas, graph.loop_info)
        return o

    def close(self):
        if getattr(self, "_h", None):
            lib().ipc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_option(self, name: str, value: float):
        _chk(lib().ipc_set_option(self._h, name.encode(), float(value)))

    # ---- reference API -------------------------------------------------------------------------
    def agreementCheck(self, edge):
        """bool IPC::agreementCheck(EDGE*) — src/consensus.cpp:42-75. Returns (accepted, CheckInfo)."""
        frm, to, meas, info = edge
        m, i = _f64(meas), _f64(info)
        acc = C.c_int(0)
        ci = CheckInfo()
        _chk(lib().ipc_agreement_check(self._h, int(frm), int(to), _p(m), _p(i), C.byref(acc), C.byref(ci)))
        return bool(acc.value), ci

    def agreementCheckStream(self, frm, to, meas, info):
        """The candidate loop of simulating_incremental_data (src/simulation.cpp:34-47) over the given candidates in order:
        same results as calling agreementCheck one by one. Returns (accepted[bool], info[INFO_DTYPE])."""
        f, t, m, i = _i32(frm), _i32(to), _f64(meas), _f64(info)
        n = f.shape[0]
        acc = np.zeros(n, dtype=np.int32)
        out = np.zeros(n, dtype=INFO_DTYPE)
        _chk(lib().ipc_agreement_check_stream(self._h, n, _p(f), _p(t), _p(m), _p(i), _p(acc), _p(out)))
        return acc.astype(bool), out

    def removeEdgeFromCnS(self, edge) -> bool:
        r = C.c_int(0)
        _chk(lib().ipc_remove_edge(self._h, int(edge[0]), int(edge[1]), C.byref(r)))
        return bool(r.value)

    def addEdgeToCnS(self, edge) -> None:
        frm, to, meas, info = edge
        m, i = _f64(meas), _f64(info)
        _chk(lib().ipc_add_edge(self._h, int(frm), int(to), _p(m), _p(i)))

    def getMaxConsensusSet(self) -> np.ndarray:
        n = C.c_int(0)
        _chk(lib().ipc_consensus_size(self._h, C.byref(n)))
        out = np.zeros((n.value, 2), dtype=np.int32)
        if n.value:
            _chk(lib().ipc_get_consensus(self._h, _p(out), n.value))
        return out

    def poses(self) -> np.ndarray:
        out = np.zeros((self.n_poses, self.mw), dtype=np.float64)
        _chk(lib().ipc_get_poses(self._h, _p(out)))
        return out

    def final_optimize(self, max_iterations: int = 1000):
        """Final full-graph optimisation of simulating_incremental_data (src/simulation.cpp:50-65). Returns (chi2, iterations)."""
        chi2, it = C.c_double(0), C.c_int(0)
        _chk(lib().ipc_final_optimize(self._h, int(max_iterations), C.byref(chi2), C.byref(it)))
        return chi2.value, it.value

    def stream_profile(self, reset: bool = False) -> dict:
        o = np.zeros(16)
        _chk(lib().ipc_stream_profile(self._h, _p(o), int(reset)))
        names = ["setup", "assemble", "factor", "back_substitute", "gn_step", "steepest_descent", "trial_states", "commit"]
        d = {f"{k}_s": float(o[i]) for i, k in enumerate(names)}
        d.update(calls=int(o[8]), factorisations=int(o[9]), trial_states=int(o[10]))
        d.update({f"factor_{k}_s": float(o[11 + i]) for i, k in enumerate(["diag", "panel", "barrier1", "update", "barrier2"])})
        return d

    # ---- batched path --------------------------------------------------------------------------
    def set_candidates(self, frm, to, meas, info):
        f, t, m, i = _i32(frm), _i32(to), _f64(meas), _f64(info)
        _chk(lib().ipc_set_candidates(self._h, f.shape[0], _p(f), _p(t), _p(m), _p(i)))
        self.n_candidates = int(f.shape[0])

    def check_batch(self, member, cand, want_info: bool = True):
        """Independent checks (host buffers, end to end). Returns (accepted[bool], info[INFO_DTYPE] | None)."""
        mb, cd = _i32(member), _i32(cand)
        n = cd.shape[0]
        bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        info = np.zeros(n, dtype=INFO_DTYPE) if want_info else None
        _chk(lib().ipc_check_batch(self._h, n, _p(mb), _p(cd), _p(bits), _p(info)))
        acc = ((bits[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
        return acc, info

    def check_batch_dev(self, n, member_ptr, cand_ptr, bits_ptr, info_ptr=None, stream=None):
        """Device-resident variant: raw device pointers (ints), enqueued on `stream` (a cudaStream_t value)."""
        _chk(lib().ipc_check_batch_dev(self._h, int(n), C.c_void_p(member_ptr), C.c_void_p(cand_ptr), C.c_void_p(bits_ptr),
                                       C.c_void_p(info_ptr) if info_ptr else None, C.c_void_p(stream) if stream else None))

    def last_batch_stats(self):
        sl, sk, nl = C.c_int64(0), C.c_int64(0), C.c_int(0)
        _chk(lib().ipc_last_batch_stats(self._h, C.byref(sl), C.byref(sk), C.byref(nl)))
        return sl.value, sk.value, nl.value

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        _chk(lib().ipc_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def consistency_matrix(self, sharded: bool = False):
        """Rows of the N_c x N_c consistency matrix (time order). sharded=True deals the solved checks over the ranks of the
        handle's communicator (comm_init) and all-gathers the verdict words: every rank gets the same rows."""
        n = self.n_candidates
        words = (n + 31) // 32
        rows = np.zeros((n, words), dtype=np.uint32)
        order = np.zeros(n, dtype=np.int32)
        solved = C.c_int64(0)
        fn = lib().ipc_consistency_matrix_sharded if sharded else lib().ipc_consistency_matrix
        _chk(fn(self._h, _p(rows), _p(order), C.byref(solved)))
        return rows, order, solved.value

    # ---- multi-GPU: one IPC object per GPU / process, NCCL all-gather behind the C ABI -----------
    def comm_init(self, dist=None, rank: int = 0, world: int = 1, uid: bytes | None = None):
        """Create the handle's NCCL communicator. With a torch.distributed module the id made on rank 0 travels by
        broadcast_object_list (any backend); otherwise pass the 128-byte `uid` from comm_unique_id() yourself."""
        if dist is not None:
            rank, world = dist.get_rank(), dist.get_world_size()
            box = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            uid = box[0]
        if uid is None or len(uid) != 128:
            raise IpcError("comm_init needs a torch.distributed module or a 128-byte id")
        buf = (C.c_ubyte * 128).from_buffer_copy(uid)
        _chk(lib().ipc_comm_init(self._h, buf, int(rank), int(world)))

    def comm_info(self):
        r, w, n = C.c_int(0), C.c_int(1), C.c_int64(0)
        _chk(lib().ipc_comm_info(self._h, C.byref(r), C.byref(w), C.byref(n)))
        return r.value, w.value, n.value

    def check_batch_sharded(self, member, cand, words_per_rank: int, out=None):
        """This rank's shard (host buffers) + ONE all-gather: returns uint32 [world, words_per_rank], same on every rank."""
        mb, cd = _i32(member), _i32(cand)
        _, world, _ = self.comm_info()
        if out is None:
            out = np.zeros((world, words_per_rank), dtype=np.uint32)
        _chk(lib().ipc_check_batch_sharded(self._h, cd.shape[0], _p(mb), _p(cd), int(words_per_rank), _p(out)))
        return out

    def check_batch_sharded_dev(self, n_local, member_ptr, cand_ptr, words_per_rank, bits_all_ptr, stream=None):
        _chk(lib().ipc_check_batch_sharded_dev(self._h, int(n_local), C.c_void_p(member_ptr), C.c_void_p(cand_ptr), int(words_per_rank),
                                               C.c_void_p(bits_all_ptr), C.c_void_p(stream) if stream else None))

    def greedy_consensus(self, rows_bits) -> np.ndarray:
        rows = np.ascontiguousarray(rows_bits, dtype=np.uint32)
        n = rows.shape[0]
        out = np.zeros(n, dtype=np.uint8)
        _chk(lib().ipc_greedy_consensus(self._h, _p(rows), n, _p(out)))
        return out.astype(bool)


def comm_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    _chk(lib().ipc_comm_unique_id(buf))
    return bytes(buf)


def pair_checks(graph, order=None):
    """Enumerate the solved checks of the consistency matrix: diagonal (fast) + overlapping pairs i<j in time
    order. Returns (member, cand) int32 arrays of loop indices (member = -1 on the diagonal)."""
    order = graph.time_order() if order is None else np.asarray(order)
    a = np.minimum(graph.loop_from, graph.loop_to)[order]
    b = np.maximum(graph.loop_from, graph.loop_to)[order]
    n = len(order)
    mem = [np.full(n, -1, dtype=np.int32)]
    cnd = [order.astype(np.int32)]
    for j in range(1, n):
        ov = (np.minimum(b[:j], b[j]) - np.maximum(a[:j], a[j])) > 0
        idx = np.nonzero(ov)[0]
        if idx.size:
            mem.append(order[idx].astype(np.int32))
            cnd.append(np.full(idx.size, order[j], dtype=np.int32))
    return np.concatenate(mem), np.concatenate(cnd)


def checks_to_csr(member, cand):
    """(member, cand) check list -> CSR loop lists (all but the last entry of a row are consensus members,
    the last is the candidate): the format the oracle's batch entry point takes."""
    member, cand = np.asarray(member), np.asarray(cand)
    k = 1 + (member >= 0).astype(np.int64)
    ptr = np.zeros(len(cand) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum(k)
    idx = np.zeros(int(ptr[-1]), dtype=np.int32)
    idx[ptr[1:] - 1] = cand
    two = member >= 0
    idx[ptr[:-1][two]] = member[two]
    return ptr, idx
