// ipc_b200.hpp — header-only C++ face of libipc_b200.so with the shape of the reference's class
//
//     template <class EDGE, class VERTEX> class IPC        /root/reference/include/ipc/consensus.hpp:5-33
//         IPC(g2o::SparseOptimizer& open_loop_problem, const Config& cfg);      :9    src/consensus.cpp:9-33
//         ~IPC();                                                                :10   src/consensus.cpp:35-40
//         bool agreementCheck(EDGE* loop_candidate);                             :12   src/consensus.cpp:42-75
//         bool removeEdgeFromCnS(EDGE* edge);                                    :13   src/consensus.cpp:77-98
//         void addEdgeToCnS(EDGE* edge);                                         :14   src/consensus.cpp:100-121
//         const std::vector<EDGE*>& getMaxConsensusSet() const;                  :16
//
// so a `simulating_incremental_data`-style caller (src/simulation.cpp:8-108) ports by swapping the type: same method names, same
// argument meaning, same return values. g2o objects are replaced by flat values (ipc_b200::Edge: vertex ids, measurement
// (x y theta | x y z qx qy qz qw), information row-major d x d) because the library does not link g2o; INTEGRATION.md shows the
// adapter that fills them from g2o::EdgeSE2 / EdgeSE3. Where the reference has undefined behaviour (ids outside the graph: null
// dereference in src/consensus_utils.cpp:32-40) this class throws ipc_b200::Error carrying ipc_last_error(). There is no CPU
// path: constructing an IPC without a CUDA device throws.
//
// Beyond the reference's six members: the batched pair checks / consistency matrix / greedy growth of the north-star path, the
// whole candidate loop in one call, the final optimisation and the estimates, the multi-GPU communicator — thin calls into
// include/ipc_b200.h, one per entry point.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ipc_b200.h"

namespace ipc_b200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
    if (rc != IPC_OK) throw Error(rc, ipc_last_error());
}

// One relative-pose edge as flat values. DIM 2: meas = (x, y, theta), info 3 x 3; DIM 3: meas = (x, y, z, qx, qy, qz, qw), info 6 x 6.
struct Edge {
    int from = 0, to = 0;
    std::vector<double> meas, info;
};

// The fields of the reference's Config that the IPC class reads (include/ipc/utils.hpp:22-38; src/consensus.cpp:17-22).
struct Config {
    double s_factor = 1.0;
    double fast_reject_th = 0.0, slow_reject_th = 0.0;
    int fast_reject_iter_base = 0, slow_reject_iter_base = 0;
};

template <int DIM> class IPC {
    static_assert(DIM == 2 || DIM == 3, "SE(2) or SE(3)");

public:
    static constexpr int kMeas = DIM == 2 ? 3 : 7;
    static constexpr int kInfo = DIM == 2 ? 9 : 36;

    // `odometry`: the edges j -> j + 1 of the open-loop problem, any order (they are sorted by head id like src/consensus.cpp:15);
    // n_poses = odometry.size() + 1. Scales their information by cfg.s_factor and dead-reckons the estimates (src/consensus.cpp:24-32).
    IPC(const std::vector<Edge>& odometry, const Config& cfg, int device = 0) {
        const int n = (int)odometry.size() + 1;
        std::vector<double> meas((size_t)kMeas * (n - 1)), info((size_t)kInfo * (n - 1));
        std::vector<char> seen(n > 1 ? n - 1 : 0, 0);
        for (const Edge& e : odometry) {
            const int lo = e.from < e.to ? e.from : e.to;
            if (e.to != e.from + 1 || lo < 0 || lo >= n - 1 || seen[lo] || (int)e.meas.size() != kMeas || (int)e.info.size() != kInfo)
                throw Error(IPC_ERR_ARG, "IPC: the odometry must be exactly one edge j -> j + 1 per consecutive pair of vertices 0 .. N-1");
            seen[lo] = 1;
            for (int q = 0; q < kMeas; ++q) meas[(size_t)kMeas * lo + q] = e.meas[q];
            for (int q = 0; q < kInfo; ++q) info[(size_t)kInfo * lo + q] = e.info[q];
        }
        init(n, meas.data(), info.data(), cfg, device);
    }
    // Same from flat arrays already in edge order: odom_meas [n_poses - 1][3 | 7], odom_info [n_poses - 1][d * d].
    IPC(int n_poses, const double* odom_meas, const double* odom_info, const Config& cfg, int device = 0) { init(n_poses, odom_meas, odom_info, cfg, device); }
    ~IPC() { ipc_destroy(h_); }
    IPC(const IPC&) = delete;
    IPC& operator=(const IPC&) = delete;
    IPC(IPC&& o) noexcept : h_(o.h_), n_poses_(o.n_poses_), n_candidates_(o.n_candidates_) { o.h_ = nullptr; }

    // ---- the reference's members -------------------------------------------------------------------------------------------
    // true: the candidate agrees with the current consensus set; it has joined the set and the estimates behind its window were
    // dead-reckoned again. false: state untouched (src/consensus.cpp:63-71).
    bool agreementCheck(const Edge& loop_candidate, ipc_check_info* info = nullptr) {
        need(loop_candidate);
        int acc = 0;
        check(ipc_agreement_check(h_, loop_candidate.from, loop_candidate.to, loop_candidate.meas.data(), loop_candidate.info.data(), &acc, info));
        return acc != 0;
    }
    bool removeEdgeFromCnS(const Edge& edge) {
        int removed = 0;
        check(ipc_remove_edge(h_, edge.from, edge.to, &removed));
        return removed != 0;
    }
    void addEdgeToCnS(const Edge& edge) {
        need(edge);
        check(ipc_add_edge(h_, edge.from, edge.to, edge.meas.data(), edge.info.data()));
    }
    // (from, to) of every member, in the set's order (sorted by the later vertex id, src/utils.cpp:371-377)
    std::vector<std::pair<int, int>> getMaxConsensusSet() const {
        int n = 0;
        check(ipc_consensus_size(h_, &n));
        std::vector<int> ft((size_t)2 * n);
        if (n) check(ipc_get_consensus(h_, ft.data(), n));
        std::vector<std::pair<int, int>> out(n);
        for (int i = 0; i < n; ++i) out[i] = {ft[2 * i], ft[2 * i + 1]};
        return out;
    }

    // ---- the caller's loop and outputs (src/simulation.cpp) ------------------------------------------------------------------
    // for (c : candidates) agreementCheck(c) — :34-47 — in one call, same sequential semantics; returns the verdict of every candidate
    std::vector<bool> agreementCheckStream(const std::vector<Edge>& candidates, std::vector<ipc_check_info>* info = nullptr) {
        const int n = (int)candidates.size();
        Flat f = flatten(candidates);
        std::vector<int> acc(n);
        if (info) info->resize(n);
        if (n) check(ipc_agreement_check_stream(h_, n, f.from.data(), f.to.data(), f.meas.data(), f.info.data(), acc.data(), info ? info->data() : nullptr));
        return std::vector<bool>(acc.begin(), acc.end());
    }
    // the estimates simulation.cpp:93-97 writes to the trajectory file: [n_poses][3 | 7]
    std::vector<double> poses() const {
        std::vector<double> p((size_t)kMeas * n_poses_);
        check(ipc_get_poses(h_, p.data()));
        return p;
    }
    // propagateGuess + odometry information divided back by s_factor + optimize(max_iterations) with the consensus set — :50-65
    double finalOptimize(int max_iterations = 1000, int* iterations = nullptr) {
        double chi2 = 0;
        int it = 0;
        check(ipc_final_optimize(h_, max_iterations, &chi2, &it));
        if (iterations) *iterations = it;
        return chi2;
    }

    // ---- the batched path (north-star): candidate table, independent checks, consistency matrix, greedy growth ---------------
    void setCandidates(const std::vector<Edge>& loops) {
        Flat f = flatten(loops);
        check(ipc_set_candidates(h_, (int)loops.size(), f.from.data(), f.to.data(), f.meas.data(), f.info.data()));
        n_candidates_ = (int)loops.size();
    }
    int candidates() const { return n_candidates_; }
    // check i: fresh IPC; member[i] >= 0 ? addEdgeToCnS(candidate member[i]) : nothing; agreementCheck(candidate cand[i]). Bit i of the
    // returned words is the verdict.
    std::vector<uint32_t> checkBatch(const std::vector<int>& member, const std::vector<int>& cand, std::vector<ipc_check_info>* info = nullptr) {
        if (member.size() != cand.size()) throw Error(IPC_ERR_ARG, "checkBatch: member and cand differ in length");
        const int n = (int)cand.size();
        std::vector<uint32_t> bits((n + 31) / 32);
        if (info) info->resize(n);
        check(ipc_check_batch(h_, n, member.data(), cand.data(), bits.data(), info ? info->data() : nullptr));
        return bits;
    }
    struct Matrix {
        int n = 0, words = 0;
        std::vector<uint32_t> rows;   // [n][words], symmetric, in time order
        std::vector<int> order;       // candidate index of row / column k
        int64_t solved = 0;           // checks actually solved (diagonal + overlapping pairs)
        bool at(int i, int j) const { return (rows[(size_t)i * words + (j >> 5)] >> (j & 31)) & 1u; }
    };
    // all candidates against each other; with a communicator (commInit) the solves are dealt over its ranks and every rank gets the rows
    Matrix consistencyMatrix(bool sharded = false) {
        Matrix m;
        m.n = n_candidates_; m.words = (m.n + 31) / 32;
        m.rows.assign((size_t)m.n * m.words, 0u); m.order.assign(m.n, 0);
        if (m.n) check((sharded ? ipc_consistency_matrix_sharded : ipc_consistency_matrix)(h_, m.rows.data(), m.order.data(), &m.solved));
        return m;
    }
    // greedy growth over the matrix in its order: a candidate joins iff it is consistent with every member so far
    std::vector<unsigned char> greedyConsensus(const Matrix& m) {
        std::vector<unsigned char> in(m.n);
        if (m.n) check(ipc_greedy_consensus(h_, m.rows.data(), m.n, in.data()));
        return in;
    }

    // ---- multi-GPU: one IPC per device / process; rank 0 makes the id and hands it to the others by any host channel ----------
    static std::vector<unsigned char> commUniqueId() {
        std::vector<unsigned char> id(IPC_COMM_ID_BYTES);
        check(ipc_comm_unique_id(id.data()));
        return id;
    }
    void commInit(const std::vector<unsigned char>& id, int rank, int world) {
        if ((int)id.size() != IPC_COMM_ID_BYTES) throw Error(IPC_ERR_ARG, "commInit: the id has IPC_COMM_ID_BYTES bytes");
        check(ipc_comm_init(h_, id.data(), rank, world));
    }

    void setOption(const char* name, double value) { check(ipc_set_option(h_, name, value)); }
    int poseCount() const { return n_poses_; }
    ipc_handle* handle() const { return h_; }

private:
    struct Flat { std::vector<int> from, to; std::vector<double> meas, info; };
    static void need(const Edge& e) {
        if ((int)e.meas.size() != kMeas || (int)e.info.size() != kInfo) throw Error(IPC_ERR_ARG, "edge: measurement / information of the wrong size for this dimension");
    }
    static Flat flatten(const std::vector<Edge>& v) {
        Flat f;
        f.from.reserve(v.size()); f.to.reserve(v.size()); f.meas.reserve(v.size() * kMeas); f.info.reserve(v.size() * kInfo);
        for (const Edge& e : v) {
            need(e);
            f.from.push_back(e.from); f.to.push_back(e.to);
            f.meas.insert(f.meas.end(), e.meas.begin(), e.meas.end());
            f.info.insert(f.info.end(), e.info.begin(), e.info.end());
        }
        return f;
    }
    void init(int n_poses, const double* meas, const double* info, const Config& cfg, int device) {
        ipc_config c{cfg.s_factor, cfg.fast_reject_th, cfg.slow_reject_th, cfg.fast_reject_iter_base, cfg.slow_reject_iter_base};
        check(ipc_create(DIM, n_poses, meas, info, &c, device, &h_));
        n_poses_ = n_poses;
    }
    ipc_handle* h_ = nullptr;
    int n_poses_ = 0, n_candidates_ = 0;
};

using IPC2D = IPC<2>;   // IPC<g2o::EdgeSE2, g2o::VertexSE2>, src/consensus.cpp:174
using IPC3D = IPC<3>;   // IPC<g2o::EdgeSE3, g2o::VertexSE3>, src/consensus.cpp:175

}  // namespace ipc_b200
