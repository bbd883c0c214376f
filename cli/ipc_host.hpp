// ipc_host.hpp — host side of the ipc_tester CLIs: flat YAML config, g2o-format graph files, trajectory / .PR output and the
// incremental simulation loop, re-hosted over the C ABI (include/ipc_b200.h). Plain C++17, no third-party dependency
// (the reference uses yaml-cpp and the g2o parser; neither is available offline).
//
// Reference sites mirrored (under /root/reference):
//   struct Config / readConfig            include/ipc/utils.hpp:22-38, src/utils.cpp:316-337
//   setProblem (load)                     src/utils.cpp:95-126        (optimizer.load: VERTEX_SE2 / EDGE_SE2 / *_SE3:QUAT / FIX)
//   splitProblemConstraints               src/utils.cpp:172-189       (|id1 - id0| == 1 -> odometry, else loop; FILE order)
//   simulating_incremental_data           src/simulation.cpp:8-108
//   writeVertex / readSolutionFile        src/utils.cpp:239-282
//   graph_fixer                           examples/graph_fixer.cpp:36-54
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../include/ipc_b200.h"

namespace ipc_host {

// ---- Config (include/ipc/utils.hpp:22-38) -------------------------------------------------------------------------
struct Config {
    std::string name, dataset, ground_truth, output;
    double s_factor = 1.0;
    int visualize = 0;
    int canonic_inliers = 0;
    double fast_reject_th = 0, slow_reject_th = 0;
    int fast_reject_iter_base = 0, slow_reject_iter_base = 0;
    int use_best_k_buddies = 0, k_buddies = 0, use_recovery = 0;
};

inline std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
inline std::string unquote(std::string v) {
    if (v.size() >= 2 && ((v.front() == '"' && v.back() == '"') || (v.front() == '\'' && v.back() == '\''))) v = v.substr(1, v.size() - 2);
    return v;
}
inline bool parse_bool(const std::string& v) { std::string l = v; for (auto& c : l) c = (char)tolower(c); return l == "true" || l == "1" || l == "yes" || l == "on"; }

// flat "key: value" YAML (all the reference's cfg/*.yaml files are flat scalars). Every key readConfig reads is required:
// a missing one throws, like yaml-cpp's operator[] / as<T>() does in the reference.
inline Config readConfig(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open config file " + path);
    std::map<std::string, std::string> kv;
    std::string line;
    while (std::getline(f, line)) {
        size_t h = line.find('#');
        if (h != std::string::npos) line = line.substr(0, h);
        size_t c = line.find(':');
        if (c == std::string::npos) continue;
        std::string k = trim(line.substr(0, c)), v = unquote(trim(line.substr(c + 1)));
        if (!k.empty()) kv[k] = v;
    }
    auto need = [&](const char* k) -> const std::string& {
        auto it = kv.find(k);
        if (it == kv.end()) throw std::runtime_error(std::string("config key missing: ") + k);
        return it->second;
    };
    Config c;
    c.name = need("name"); c.dataset = need("dataset"); c.ground_truth = need("ground_truth"); c.output = need("output");
    c.s_factor = std::stod(need("s_factor")); c.visualize = parse_bool(need("visualize"));
    c.canonic_inliers = std::stoi(need("canonic_inliers"));
    c.fast_reject_th = std::stod(need("fast_reject_th")); c.fast_reject_iter_base = std::stoi(need("fast_reject_iter_base"));
    c.slow_reject_th = std::stod(need("slow_reject_th")); c.slow_reject_iter_base = std::stoi(need("slow_reject_iter_base"));
    c.use_best_k_buddies = parse_bool(need("use_best_k_buddies")); c.k_buddies = std::stoi(need("k_buddies")); c.use_recovery = parse_bool(need("use_recovery"));
    return c;
}

// ---- graph (g2o text format) --------------------------------------------------------------------------------------
struct EdgeRec { int from, to; std::vector<double> meas, info; };   // meas 3 | 7, info d*d row-major full symmetric
struct Graph {
    int dim = 2;
    std::vector<int> vertex_ids;
    std::vector<std::vector<double>> vertex_est;
    std::vector<EdgeRec> edges;        // FILE order
    std::vector<int> fixed;
};

// optimizer.load (src/utils.cpp:114, g2o core/optimizable_graph.cpp::load) on a fast path: the file is read in one piece, cut
// into chunks at line boundaries, and the chunks are tokenised by worker threads with strtod / strtol straight from the buffer
// (no per-line stream objects); the pieces are concatenated in FILE order, so candidate order and labels do not depend on the
// thread count. Unknown tags are skipped, like g2o's loader (it prints a warning); FIX lines are recorded.
namespace detail {
struct Piece { std::vector<int> vertex_ids; std::vector<std::vector<double>> vertex_est; std::vector<EdgeRec> edges; std::vector<int> fixed; std::string err; };
inline const char* skip_ws(const char* p, const char* e) { while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; return p; }
inline bool rd_int(const char*& p, const char* e, int& v) {
    p = skip_ws(p, e); if (p >= e) return false;
    char* q = nullptr; const long x = std::strtol(p, &q, 10);
    if (q == p) return false;
    v = (int)x; p = q; return true;
}
inline bool rd_dbl(const char*& p, const char* e, double& v) {
    p = skip_ws(p, e); if (p >= e) return false;
    char* q = nullptr; const double x = std::strtod(p, &q);
    if (q == p) return false;
    v = x; p = q; return true;
}
inline void parse_chunk(const char* b, const char* e, int dim, Piece& out) {
    const char* vtag = dim == 2 ? "VERTEX_SE2" : "VERTEX_SE3:QUAT"; const char* etag = dim == 2 ? "EDGE_SE2" : "EDGE_SE3:QUAT";
    const size_t vl = std::strlen(vtag), el = std::strlen(etag);
    const int mw = dim == 2 ? 3 : 7, d = dim == 2 ? 3 : 6;
    std::vector<double> up((size_t)d * (d + 1) / 2);
    while (b < e) {
        const char* nl = (const char*)std::memchr(b, '\n', (size_t)(e - b));
        const char* le = nl ? nl : e;                       // the buffer is NUL-terminated past e, so strtod cannot run away
        const char* p = skip_ws(b, le);
        const char* t = p; while (t < le && *t != ' ' && *t != '\t' && *t != '\r') ++t;
        const size_t tl = (size_t)(t - p);
        bool bad = false;
        if (tl == vl && std::memcmp(p, vtag, vl) == 0) {
            int id; const char* q = t; std::vector<double> est(mw);
            bad = !rd_int(q, le, id); for (int c = 0; c < mw && !bad; ++c) bad = !rd_dbl(q, le, est[c]);
            if (!bad) { out.vertex_ids.push_back(id); out.vertex_est.push_back(std::move(est)); }
        } else if (tl == el && std::memcmp(p, etag, el) == 0) {
            EdgeRec r; const char* q = t; r.meas.resize(mw);
            bad = !rd_int(q, le, r.from) || !rd_int(q, le, r.to);
            for (int c = 0; c < mw && !bad; ++c) bad = !rd_dbl(q, le, r.meas[c]);
            for (size_t c = 0; c < up.size() && !bad; ++c) bad = !rd_dbl(q, le, up[c]);       // upper triangle, row-major
            if (!bad) {
                r.info.assign((size_t)d * d, 0.0);
                int k = 0;
                for (int i = 0; i < d; ++i) for (int j = i; j < d; ++j) { r.info[i * d + j] = up[k]; r.info[j * d + i] = up[k]; ++k; }
                out.edges.push_back(std::move(r));
            }
        } else if (tl == 3 && std::memcmp(p, "FIX", 3) == 0) { const char* q = t; int id; while (rd_int(q, le, id)) out.fixed.push_back(id); }
        if (bad && out.err.empty()) out.err = std::string("malformed line: ") + std::string(b, le);
        b = nl ? nl + 1 : e;
    }
}
}  // namespace detail

inline Graph loadG2O(const std::string& path, int dim, int n_threads = 0) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open dataset " + path);
    f.seekg(0, std::ios::end);
    const std::streamoff len = f.tellg();
    f.seekg(0);
    std::string buf((size_t)len, '\0');                    // std::string keeps a NUL behind the last byte
    if (len > 0) f.read(&buf[0], len);
    if (n_threads <= 0) { n_threads = (int)std::thread::hardware_concurrency(); if (n_threads <= 0) n_threads = 1; if (n_threads > 16) n_threads = 16; }
    const size_t min_chunk = 1 << 16;
    int nt = (int)std::min<size_t>((size_t)n_threads, buf.size() / min_chunk + 1);
    std::vector<size_t> cut(nt + 1, buf.size());
    cut[0] = 0;
    for (int i = 1; i < nt; ++i) {                          // chunk i starts behind the first newline at or after its even split
        size_t p = buf.size() / nt * i;
        if (p < cut[i - 1]) p = cut[i - 1];
        const size_t q = buf.find('\n', p);
        cut[i] = q == std::string::npos ? buf.size() : q + 1;
    }
    std::vector<detail::Piece> pieces(nt);
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back([&, i] { detail::parse_chunk(buf.data() + cut[i], buf.data() + cut[i + 1], dim, pieces[i]); });
    detail::parse_chunk(buf.data() + cut[0], buf.data() + cut[1], dim, pieces[0]);
    for (auto& t : th) t.join();
    Graph g; g.dim = dim;
    for (auto& p : pieces) {
        if (!p.err.empty()) throw std::runtime_error(p.err);
        g.vertex_ids.insert(g.vertex_ids.end(), p.vertex_ids.begin(), p.vertex_ids.end());
        for (auto& v : p.vertex_est) g.vertex_est.push_back(std::move(v));
        for (auto& e : p.edges) g.edges.push_back(std::move(e));
        g.fixed.insert(g.fixed.end(), p.fixed.begin(), p.fixed.end());
    }
    return g;
}

// optimizer.save (examples/graph_fixer.cpp:54): vertices in id order, FIX lines, then the edges (FILE order here; g2o walks a
// pointer-ordered set), information as the row-major upper triangle, 17 significant digits so a reload is bit exact.
inline void saveG2O(const Graph& g, const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + path);
    const char* vtag = g.dim == 2 ? "VERTEX_SE2" : "VERTEX_SE3:QUAT"; const char* etag = g.dim == 2 ? "EDGE_SE2" : "EDGE_SE3:QUAT";
    const int d = g.dim == 2 ? 3 : 6;
    std::vector<size_t> vo(g.vertex_ids.size());
    for (size_t i = 0; i < vo.size(); ++i) vo[i] = i;
    std::stable_sort(vo.begin(), vo.end(), [&](size_t a, size_t b) { return g.vertex_ids[a] < g.vertex_ids[b]; });
    for (size_t i : vo) {
        std::fprintf(f, "%s %d", vtag, g.vertex_ids[i]);
        for (double x : g.vertex_est[i]) std::fprintf(f, " %.17g", x);
        std::fputc('\n', f);
    }
    for (int id : g.fixed) std::fprintf(f, "FIX %d\n", id);
    for (const auto& e : g.edges) {
        std::fprintf(f, "%s %d %d", etag, e.from, e.to);
        for (double x : e.meas) std::fprintf(f, " %.17g", x);
        for (int r = 0; r < d; ++r) for (int c = r; c < d; ++c) std::fprintf(f, " %.17g", e.info[(size_t)r * d + c]);
        std::fputc('\n', f);
    }
    std::fclose(f);
}

// measurement().inverse() of a relative pose: SE(2) (x y theta) / SE(3) (t, unit quaternion qx qy qz qw)
inline std::vector<double> inverseMeasurement(const std::vector<double>& m) {
    if (m.size() == 3) {
        const double c = std::cos(m[2]), s = std::sin(m[2]);
        double th = -m[2];
        th -= 6.283185307179586476925 * std::floor((th + 3.14159265358979323846) / 6.283185307179586476925);   // normalize_theta: [-pi, pi)
        return {-(c * m[0] + s * m[1]), -(-s * m[0] + c * m[1]), th};
    }
    double qx = m[3], qy = m[4], qz = m[5], qw = m[6];
    const double n = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    qx /= n; qy /= n; qz /= n; qw /= n;
    // R^T t with R from the unit quaternion; t' = -R^T t, q' = conjugate
    const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw), 2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz),
                         2 * (qy * qz - qx * qw), 2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
    const double tx = -(R[0] * m[0] + R[3] * m[1] + R[6] * m[2]), ty = -(R[1] * m[0] + R[4] * m[1] + R[7] * m[2]), tz = -(R[2] * m[0] + R[5] * m[1] + R[8] * m[2]);
    return {tx, ty, tz, -qx, -qy, -qz, qw};
}

// examples/graph_fixer.cpp:36-52: every loop edge (|from - to| != 1) stored as to < from is re-oriented: vertices swapped,
// measurement inverted, information copied UNCHANGED (the reference does not transport it to the other frame; kept as is).
// Returns the number of edges that were flipped.
inline int fixGraph(Graph& g) {
    int flipped = 0;
    for (auto& e : g.edges) {
        if (std::abs(e.to - e.from) == 1 || e.from < e.to) continue;
        std::swap(e.from, e.to);
        e.meas = inverseMeasurement(e.meas);
        ++flipped;
    }
    return flipped;
}

// splitProblemConstraints (src/utils.cpp:172-189) + getProblemOdom ordering (src/consensus.cpp:15): odometry j -> j+1
struct Problem {
    int dim = 2, n_poses = 0;
    std::vector<double> odom_meas, odom_info;     // [n-1][mw], [n-1][d*d]
    std::vector<EdgeRec> loops;                   // file order
};
inline Problem splitProblem(const Graph& g) {
    Problem p; p.dim = g.dim;
    const int mw = g.dim == 2 ? 3 : 7, d = g.dim == 2 ? 3 : 6;
    int n = 0;
    for (int id : g.vertex_ids) n = std::max(n, id + 1);
    for (const auto& e : g.edges) n = std::max(n, std::max(e.from, e.to) + 1);
    p.n_poses = n;
    std::vector<const EdgeRec*> od(n > 0 ? n - 1 : 0, nullptr);
    for (const auto& e : g.edges) {
        if (std::abs(e.to - e.from) == 1) {
            if (e.to != e.from + 1) throw std::runtime_error("odometry edge not oriented i -> i+1 (dataset not suitable, cf. cfg/3D/CUBE_params.yaml:12)");
            if (od[e.from]) throw std::runtime_error("duplicate odometry edge");
            od[e.from] = &e;
        } else p.loops.push_back(e);
    }
    for (int j = 0; j + 1 < n; ++j) {
        if (!od[j]) throw std::runtime_error("missing odometry edge " + std::to_string(j) + " -> " + std::to_string(j + 1) + " (vertex ids must be contiguous)");
        p.odom_meas.insert(p.odom_meas.end(), od[j]->meas.begin(), od[j]->meas.end());
        p.odom_info.insert(p.odom_info.end(), od[j]->info.begin(), od[j]->info.end());
    }
    (void)mw; (void)d;
    return p;
}

// readSolutionFile (src/utils.cpp:262-282): the ground truth must exist and parse even though it is unused
inline std::vector<std::vector<double>> readSolutionFile(const std::string& path, int dim) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open ground truth " + path);
    std::vector<std::vector<double>> out;
    std::string line;
    const int w = dim == 2 ? 3 : 7;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        std::vector<double> v(w);
        bool ok = true;
        for (auto& x : v) if (!(is >> x)) { ok = false; break; }
        if (ok) out.push_back(v);
    }
    return out;
}
inline void writeTrajectory(const std::string& path, const std::vector<double>& poses, int n, int w) {   // writeVertex, src/utils.cpp:239-258
    std::ofstream f(path);
    if (!f) throw std::runtime_error("cannot write " + path);
    f.precision(10);
    for (int i = 0; i < n; ++i) { for (int c = 0; c < w; ++c) f << poses[(size_t)i * w + c] << (c + 1 < w ? " " : "\n"); }
}

struct SimResult { int tp = 0, fp = 0, tn = 0, fn = 0; float precision = 0, recall = 0; double total_s = 0; int n_candidates = 0; std::vector<int> accepted; };

inline void check(int rc) { if (rc != IPC_OK) throw std::runtime_error(std::string("ipc_b200: ") + ipc_last_error()); }

// simulating_incremental_data (src/simulation.cpp:8-108) over the C ABI
inline SimResult simulate(const Config& cfg, const Problem& p, int device, bool final_pgo, bool quiet, bool one_by_one = false) {
    (void)readSolutionFile(cfg.ground_truth, p.dim);                         // :14-15
    const int n_loops = (int)p.loops.size();
    std::vector<int> order(n_loops);
    for (int i = 0; i < n_loops; ++i) order[i] = i;
    // labels: the first canonic_inliers loops (file order) are true (:24-25); candidates in time order (:26, stable)
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::max(p.loops[a].from, p.loops[a].to) < std::max(p.loops[b].from, p.loops[b].to); });
    ipc_config c{cfg.s_factor, cfg.fast_reject_th, cfg.slow_reject_th, cfg.fast_reject_iter_base, cfg.slow_reject_iter_base};
    ipc_handle* h = nullptr;
    check(ipc_create(p.dim, p.n_poses, p.odom_meas.data(), p.odom_info.data(), &c, device, &h));   // IPC ipc(problem, cfg), :28
    SimResult r; r.n_candidates = n_loops; r.accepted.assign(n_loops, 0);
    std::vector<int> acc_k(n_loops, 0);
    if (one_by_one) {
        for (int k = 0; k < n_loops; ++k) {                                  // :34-47, one ipc_agreement_check per candidate
            const EdgeRec& e = p.loops[order[k]];
            auto t0 = std::chrono::steady_clock::now();
            check(ipc_agreement_check(h, e.from, e.to, e.meas.data(), e.info.data(), &acc_k[k], nullptr));
            r.total_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (!quiet && (k % 50 == 0 || k + 1 == n_loops)) { std::fprintf(stderr, "\r[%d / %d]", k + 1, n_loops); std::fflush(stderr); }
        }
        if (!quiet) std::fprintf(stderr, "\n");
    } else {                                                                 // :34-47, the whole loop in one call (same results)
        const int w = p.dim == 2 ? 3 : 7, dd = p.dim == 2 ? 9 : 36;
        std::vector<int> from(n_loops), to(n_loops);
        std::vector<double> meas((size_t)n_loops * w), info((size_t)n_loops * dd);
        for (int k = 0; k < n_loops; ++k) {
            const EdgeRec& e = p.loops[order[k]];
            from[k] = e.from; to[k] = e.to;
            std::copy(e.meas.begin(), e.meas.end(), meas.begin() + (size_t)k * w);
            std::copy(e.info.begin(), e.info.end(), info.begin() + (size_t)k * dd);
        }
        auto t0 = std::chrono::steady_clock::now();
        check(ipc_agreement_check_stream(h, n_loops, from.data(), to.data(), meas.data(), info.data(), acc_k.data(), nullptr));
        r.total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    for (int k = 0; k < n_loops; ++k) {
        const int acc = acc_k[k];
        r.accepted[order[k]] = acc;
        const bool truth = order[k] < cfg.canonic_inliers;
        if (acc && truth) ++r.tp; else if (acc && !truth) ++r.fp; else if (!acc && truth) ++r.fn; else ++r.tn;
    }
    r.precision = (r.tp + r.fp) > 0 ? (float)r.tp / (float)(r.tp + r.fp) : 0.f;   // float, :80-81
    r.recall = (r.tp + r.fn) > 0 ? (float)r.tp / (float)(r.tp + r.fn) : 0.f;
    const int w = p.dim == 2 ? 3 : 7;
    std::vector<double> poses((size_t)p.n_poses * w);
    if (final_pgo) {
        double chi2 = 0; int iters = 0;
        check(ipc_final_optimize(h, 1000, &chi2, &iters));                   // :50-65
        if (!quiet) std::cout << "Final optimisation: chi2 = " << chi2 << " after " << iters << " iterations\n";
    }
    check(ipc_get_poses(h, poses.data()));
    writeTrajectory(cfg.output, poses, p.n_poses, w);                        // :91-98
    const std::string pr = cfg.output.substr(0, cfg.output.size() >= 3 ? cfg.output.size() - 3 : 0) + "PR";   // :101
    std::ofstream f(pr);
    f << r.precision << " " << r.recall << "\n" << r.total_s << " " << (n_loops ? r.total_s / n_loops : 0.0) << "\n";   // :103-104
    ipc_destroy(h);
    return r;
}


// --matrix [--gpus N]: the batch path instead of the sequential stream — the N_c x N_c pairwise consistency matrix (fast checks on the
// diagonal, overlapping pairs off it, time order of src/simulation.cpp:26) and the greedy consensus growth over it. With N > 1 one handle
// per GPU runs on its own host thread; the solved checks are dealt over the GPUs and ONE NCCL all-gather of the packed verdict words (behind
// the C ABI: ipc_comm_init / ipc_consistency_matrix_sharded) gives every rank the same rows. Labels and precision / recall as in
// simulating_incremental_data (:24-25, :80-81); writes <output minus 3 chars>PR like the stream does.
struct MatrixResult { int tp = 0, fp = 0, tn = 0, fn = 0; float precision = 0, recall = 0; double total_s = 0; long long solved = 0; int n_candidates = 0; int gpus = 1;
                      std::vector<int> accepted; };
inline MatrixResult consistency_matrix_run(const Config& cfg, const Problem& p, int gpus, int first_device, bool quiet) {
    const int n = (int)p.loops.size(), w = p.dim == 2 ? 3 : 7, dd = p.dim == 2 ? 9 : 36, words = (n + 31) / 32;
    std::vector<int> from(n), to(n);
    std::vector<double> meas((size_t)n * w), info((size_t)n * dd);
    for (int k = 0; k < n; ++k) {
        const EdgeRec& e = p.loops[k];
        from[k] = e.from; to[k] = e.to;
        std::copy(e.meas.begin(), e.meas.end(), meas.begin() + (size_t)k * w);
        std::copy(e.info.begin(), e.info.end(), info.begin() + (size_t)k * dd);
    }
    ipc_config c{cfg.s_factor, cfg.fast_reject_th, cfg.slow_reject_th, cfg.fast_reject_iter_base, cfg.slow_reject_iter_base};
    unsigned char id[IPC_COMM_ID_BYTES] = {0};
    if (gpus > 1) check(ipc_comm_unique_id(id));
    std::vector<std::vector<uint32_t>> rows(gpus);
    std::vector<std::vector<int>> order(gpus, std::vector<int>(n));
    std::vector<std::vector<unsigned char>> in_set(gpus, std::vector<unsigned char>(n, 0));
    std::vector<long long> solved(gpus, 0);
    std::vector<std::string> err(gpus);
    std::vector<double> secs(gpus, 0.0);
    auto work = [&](int r) {
        ipc_handle* h = nullptr;
        auto ok = [&](int rc) { if (rc != IPC_OK) { err[r] = ipc_last_error(); return false; } return true; };
        do {
            if (!ok(ipc_create(p.dim, p.n_poses, p.odom_meas.data(), p.odom_info.data(), &c, first_device + r, &h))) break;
            if (!ok(ipc_set_candidates(h, n, from.data(), to.data(), meas.data(), info.data()))) break;
            if (gpus > 1 && !ok(ipc_comm_init(h, id, r, gpus))) break;
            rows[r].assign((size_t)n * words, 0u);
            int64_t ns = 0;
            auto t0 = std::chrono::steady_clock::now();
            if (!ok(gpus > 1 ? ipc_consistency_matrix_sharded(h, rows[r].data(), order[r].data(), &ns)
                             : ipc_consistency_matrix(h, rows[r].data(), order[r].data(), &ns))) break;
            if (!ok(ipc_greedy_consensus(h, rows[r].data(), n, in_set[r].data()))) break;
            secs[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            solved[r] = ns;
        } while (false);
        if (h) ipc_destroy(h);
    };
    std::vector<std::thread> th;
    for (int r = 1; r < gpus; ++r) th.emplace_back(work, r);
    work(0);
    for (auto& t : th) t.join();
    for (int r = 0; r < gpus; ++r) if (!err[r].empty()) throw std::runtime_error("ipc_b200 (GPU " + std::to_string(first_device + r) + "): " + err[r]);
    for (int r = 1; r < gpus; ++r)
        if (rows[r] != rows[0] || in_set[r] != in_set[0]) throw std::runtime_error("ranks disagree on the gathered consistency matrix");
    MatrixResult m; m.n_candidates = n; m.gpus = gpus; m.solved = solved[0]; m.accepted.assign(n, 0);
    for (int r = 0; r < gpus; ++r) m.total_s = std::max(m.total_s, secs[r]);
    for (int k = 0; k < n; ++k) {                      // row k of the matrix is candidate order[k] of the file
        const int li = order[0][k]; const bool acc = in_set[0][k] != 0, truth = li < cfg.canonic_inliers;
        m.accepted[li] = acc;
        if (acc && truth) ++m.tp; else if (acc && !truth) ++m.fp; else if (!acc && truth) ++m.fn; else ++m.tn;
    }
    m.precision = (m.tp + m.fp) > 0 ? (float)m.tp / (float)(m.tp + m.fp) : 0.f;
    m.recall = (m.tp + m.fn) > 0 ? (float)m.tp / (float)(m.tp + m.fn) : 0.f;
    const std::string pr = cfg.output.substr(0, cfg.output.size() >= 3 ? cfg.output.size() - 3 : 0) + "PR";
    std::ofstream f(pr);
    f << m.precision << " " << m.recall << "\n" << m.total_s << " " << (m.solved ? m.total_s / (double)m.solved : 0.0) << "\n";
    (void)quiet;
    return m;
}

inline int tester_main(int argc, char** argv, int dim) {
    std::string cfg_path; int device = 0, gpus = 1; bool parse_only = false, quiet = false, final_pgo = true, one_by_one = false, matrix = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-c" && i + 1 < argc) cfg_path = argv[++i];
        else if (a == "--device" && i + 1 < argc) device = std::stoi(argv[++i]);
        else if (a == "--parse-only") parse_only = true;
        else if (a == "--no-final-pgo") final_pgo = false;
        else if (a == "--quiet") quiet = true;
        else if (a == "--one-by-one") one_by_one = true;
        else if (a == "--matrix") matrix = true;
        else if (a == "--gpus" && i + 1 < argc) gpus = std::stoi(argv[++i]);
        else { std::cerr << "usage: " << argv[0] << " -c <config.yaml> [--device N] [--parse-only] [--no-final-pgo] [--quiet] [--one-by-one] [--matrix [--gpus N]]\n"; return 2; }
    }
    if (cfg_path.empty()) { std::cerr << "usage: " << argv[0] << " -c <config.yaml>\n"; return 2; }
    try {
        Config cfg = readConfig(cfg_path);
        Graph g = loadG2O(cfg.dataset, dim);
        Problem p = splitProblem(g);
        std::cout << "Dataset " << cfg.name << ": " << p.n_poses << " poses, " << p.n_poses - 1 << " odometry edges, " << p.loops.size() << " loop candidates ("
                  << cfg.canonic_inliers << " canonic inliers)\n";
        if (parse_only) {
            std::cout << "s_factor " << cfg.s_factor << " fast " << cfg.fast_reject_th << "/" << cfg.fast_reject_iter_base << " slow " << cfg.slow_reject_th << "/"
                      << cfg.slow_reject_iter_base << "\n";
            return 0;
        }
        if (matrix) {
            if (gpus < 1 || device + gpus > ipc_device_count()) throw std::runtime_error("--gpus " + std::to_string(gpus) + " from device " + std::to_string(device) + ": only " +
                                                                                        std::to_string(ipc_device_count()) + " CUDA device(s)");
            (void)readSolutionFile(cfg.ground_truth, p.dim);
            MatrixResult m = consistency_matrix_run(cfg, p, gpus, device, quiet);
            std::cout << "Consistency matrix: " << m.n_candidates << " candidates, " << m.solved << " solved checks on " << m.gpus << " GPU(s) in " << m.total_s << " s ("
                      << (m.total_s > 0 ? (double)m.solved / m.total_s : 0.0) << " checks/s incl. transfers)\n";
            std::cout << "TP " << m.tp << " FP " << m.fp << " TN " << m.tn << " FN " << m.fn << "\n";
            std::cout << "Precision = " << m.precision << "  Recall = " << m.recall << "\n";
            return 0;
        }
        SimResult r = simulate(cfg, p, device, final_pgo, quiet, one_by_one);
        std::cout << "TP " << r.tp << " FP " << r.fp << " TN " << r.tn << " FN " << r.fn << "\n";
        std::cout << "Precision = " << r.precision << "  Recall = " << r.recall << "\n";
        std::cout << "Total time = " << r.total_s << " s  Avg Time x test = " << (r.n_candidates ? r.total_s / r.n_candidates : 0.0) << " s\n";
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}

// examples/graph_fixer.cpp: `graph_fixer -c cfg.yaml` loads cfg.dataset (SE(3) in the reference; --dim 2 for SE(2) files),
// re-orients the loop edges and writes ./graph.g2o (or -o <path>).
inline int graph_fixer_main(int argc, char** argv) {
    std::string cfg_path, out = "graph.g2o"; int dim = 3;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-c" && i + 1 < argc) cfg_path = argv[++i];
        else if (a == "-o" && i + 1 < argc) out = argv[++i];
        else if (a == "--dim" && i + 1 < argc) dim = std::stoi(argv[++i]);
        else { std::cerr << "usage: " << argv[0] << " -c <config.yaml> [-o graph.g2o] [--dim 2|3]\n"; return 2; }
    }
    if (cfg_path.empty() || (dim != 2 && dim != 3)) { std::cerr << "usage: " << argv[0] << " -c <config.yaml> [-o graph.g2o] [--dim 2|3]\n"; return 2; }
    try {
        Config cfg = readConfig(cfg_path);
        Graph g = loadG2O(cfg.dataset, dim);
        std::cout << "Tot Edges = " << g.edges.size() << "\n" << "Tot Vertices = " << g.vertex_ids.size() << "\n";   // examples/graph_fixer.cpp:30-31
        const int n = fixGraph(g);
        saveG2O(g, out);
        std::cout << "Re-oriented " << n << " loop edges -> " << out << "\n";
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}

}  // namespace ipc_host
