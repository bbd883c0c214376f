// host_state.hpp — host mirror of the IPC object: odometry records, consensus set and the integer
// logic of /root/reference/src/consensus.cpp:77-171 (consensus-set edits, cluster discovery).
// No solver lives here: all numerical work runs on the GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <utility>
#include <vector>

namespace ipcb {

struct HostEdge {
    int from = 0, to = 0;
    std::vector<double> meas;   // 3 | 7
    std::vector<double> info;   // d*d
};

struct HostState {
    int dim = 2, d = 3, mw = 3, n = 0;
    std::vector<double> odom_meas;   // [n-1][mw]
    std::vector<double> odom_info;   // [n-1][d*d], already multiplied by s_factor (robustifyVoters, consensus_utils.cpp:123-130)
    std::vector<double> odom_info_raw;   // as given (the simulation divides s_factor back before the final optimisation, simulation.cpp:55-56)
    std::vector<HostEdge> cns;       // _max_consensus_set
    bool uniform_iso = false;        // every odometry edge carries the same information, isotropic in (x, y) and with no
    double Du[6] = {0}, Vu[6] = {0}; // x/y-theta coupling: E^T Omega E = Omega for every edge, kernels skip the per-edge loads

    static double norm_theta(double t) {
        const double pi = 3.14159265358979323846;
        if (t >= -pi && t < pi) return t;
        double m = std::floor(t / (2 * pi));
        t = t - m * 2 * pi;
        if (t >= pi) t -= 2 * pi;
        if (t < -pi) t += 2 * pi;
        return t;
    }

    bool init(int dim_, int n_poses, const double* om, const double* oi, double s_factor, std::string& err) {
        dim = dim_; d = dim == 2 ? 3 : 6; mw = dim == 2 ? 3 : 7; n = n_poses;
        odom_meas.assign(om, om + (size_t)(n - 1) * mw);
        odom_info.assign(oi, oi + (size_t)(n - 1) * d * d);
        odom_info_raw = odom_info;
        for (auto& v : odom_info) v *= s_factor;
        for (size_t i = 0; i < odom_meas.size(); ++i) if (!std::isfinite(odom_meas[i])) { err = "non-finite odometry measurement"; return false; }
        for (size_t i = 0; i < odom_info.size(); ++i) if (!std::isfinite(odom_info[i])) { err = "non-finite odometry information"; return false; }
        uniform_iso = false;
        if (dim == 2) {
            const double* W = odom_info.data();
            bool u = W[0] == W[4] && W[1] == 0 && W[3] == 0 && W[2] == 0 && W[5] == 0 && W[6] == 0 && W[7] == 0;
            for (int k = 1; u && k + 1 < n; ++k) u = std::equal(W, W + 9, W + (size_t)k * 9);
            if (u) {
                uniform_iso = true;
                Du[0] = W[0]; Du[1] = 0; Du[2] = 0; Du[3] = W[4]; Du[4] = 0; Du[5] = W[8];
                Vu[0] = 1.0 / W[0]; Vu[1] = 0; Vu[2] = 0; Vu[3] = 1.0 / W[4]; Vu[4] = 0; Vu[5] = 1.0 / W[8];
            }
        }
        return true;
    }

    // SE(2) record in the edge's relative-pose frame: residual d = r - z with r = Xi^-1 Xj as (x, y, theta),
    // chi2 = d^T D d, D = E^T (s * Omega) E, E = blockdiag(R_z^T, 1)   (e_g2o = E d, edge_se2.cpp)
    static void se2_edge_record(const double* meas, const double* info, double scale, double* z_out, double* D_out) {
        z_out[0] = meas[0]; z_out[1] = meas[1]; z_out[2] = norm_theta(meas[2]);
        double c = std::cos(z_out[2]), s = std::sin(z_out[2]);
        double E[9] = {c, s, 0, -s, c, 0, 0, 0, 1};
        double W[9], T[9];
        for (int i = 0; i < 9; ++i) W[i] = info[i] * scale;
        for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) { double a = 0; for (int k = 0; k < 3; ++k) a += W[r * 3 + k] * E[k * 3 + q]; T[r * 3 + q] = a; }
        double Dm[9];
        for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) { double a = 0; for (int k = 0; k < 3; ++k) a += E[k * 3 + r] * T[k * 3 + q]; Dm[r * 3 + q] = a; }
        D_out[0] = Dm[0]; D_out[1] = 0.5 * (Dm[1] + Dm[3]); D_out[2] = 0.5 * (Dm[2] + Dm[6]);
        D_out[3] = Dm[4]; D_out[4] = 0.5 * (Dm[5] + Dm[7]); D_out[5] = Dm[8];
    }
    static void inv_sym3_host(const double* D, double* V) {
        double c00 = D[3] * D[5] - D[4] * D[4], c01 = D[2] * D[4] - D[1] * D[5], c02 = D[1] * D[4] - D[2] * D[3];
        double id = 1.0 / (D[0] * c00 + D[1] * c01 + D[2] * c02);
        V[0] = c00 * id; V[1] = c01 * id; V[2] = c02 * id;
        V[3] = (D[0] * D[5] - D[2] * D[2]) * id; V[4] = (D[1] * D[2] - D[0] * D[4]) * id; V[5] = (D[0] * D[3] - D[1] * D[1]) * id;
    }

    // AoS layout in HBM: (zx zy zt) per edge when uniform_iso (24 B), else (zx zy zt d00 d01 d02 d11 d12 d22) (72 B)
    int odom_rec_doubles(bool uni) const { return uni ? 3 : 9; }
    void build_odom_aos(bool uni, int n_pad, std::vector<double>& rec, bool raw = false) const {
        const std::vector<double>& odom_info = raw ? odom_info_raw : this->odom_info;
        const int w = odom_rec_doubles(uni);
        rec.assign((size_t)w * n_pad, 0.0);
        if (dim != 2) { rec.clear(); return; }
        for (int k = 0; k < n_pad; ++k) {
            double* r = &rec[(size_t)k * w];
            if (k + 1 < n) {
                double z[3], D[6];
                se2_edge_record(&odom_meas[(size_t)k * 3], &odom_info[(size_t)k * 9], 1.0, z, D);
                for (int c = 0; c < 3; ++c) r[c] = z[c];
                if (!uni) for (int c = 0; c < 6; ++c) r[3 + c] = D[c];
            } else if (!uni) { r[3] = 1; r[6] = 1; r[8] = 1; }   // padding entries keep D = identity so stray reads stay finite
        }
    }

    // ---- SE(3) records: Z^-1 as (t, q = w x y z), Omega packed (21, row-major upper), V = Omega^-1 packed (21) -------------
    static bool inv_spd6(const double* A, double* Ainv) {   // Gauss-Jordan on a symmetric positive definite 6x6
        double M[6][12];
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { M[r][c] = A[r * 6 + c]; M[r][6 + c] = r == c ? 1.0 : 0.0; }
        for (int c = 0; c < 6; ++c) {
            int p = c; for (int r = c + 1; r < 6; ++r) if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
            if (M[p][c] == 0.0) return false;
            if (p != c) for (int k = 0; k < 12; ++k) std::swap(M[p][k], M[c][k]);
            double inv = 1.0 / M[c][c];
            for (int k = 0; k < 12; ++k) M[c][k] *= inv;
            for (int r = 0; r < 6; ++r) if (r != c) { double f = M[r][c]; if (f != 0.0) for (int k = 0; k < 12; ++k) M[r][k] -= f * M[c][k]; }
        }
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Ainv[r * 6 + c] = 0.5 * (M[r][6 + c] + M[c][6 + r]);
        return true;
    }
    static bool se3_edge_record(const double* meas, const double* info, double scale, double* rec /* 49 */) {
        // g2o reader: quaternion (qx qy qz qw in the file) normalised, w >= 0 (EDGE_SE3:QUAT)
        double q[4] = {meas[6], meas[3], meas[4], meas[5]};
        double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (!(n > 0)) return false;
        for (double& v : q) v /= n;
        if (q[0] < 0) for (double& v : q) v = -v;
        // Z^-1 = (R^T, -R^T t)
        const double w = q[0], x = q[1], y = q[2], z = q[3];
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                             2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
        for (int r = 0; r < 3; ++r) rec[r] = -(R[r] * meas[0] + R[3 + r] * meas[1] + R[6 + r] * meas[2]);
        rec[3] = w; rec[4] = -x; rec[5] = -y; rec[6] = -z;
        double W[36], Wi[36];
        for (int i = 0; i < 36; ++i) W[i] = info[i] * scale;
        if (!inv_spd6(W, Wi)) return false;
        int k = 0;
        for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { rec[7 + k] = 0.5 * (W[r * 6 + c] + W[c * 6 + r]); rec[28 + k] = Wi[r * 6 + c]; ++k; }
        return true;
    }
    bool build_odom_aos3(int n_pad, std::vector<double>& rec, bool raw = false) const {
        const std::vector<double>& odom_info = raw ? odom_info_raw : this->odom_info;
        rec.assign((size_t)49 * n_pad, 0.0);
        for (int k = 0; k < n_pad; ++k) {
            double* r = &rec[(size_t)k * 49];
            if (k + 1 < n) { if (!se3_edge_record(&odom_meas[(size_t)k * 7], &odom_info[(size_t)k * 36], 1.0, r)) return false; }
            else { r[3] = 1; for (int d = 0, q = 0; d < 6; ++d) { r[7 + q] = 1; r[28 + q] = 1; q += 6 - d; } }
        }
        return true;
    }

    // removeEdgeFromCnS, src/consensus.cpp:77-98
    bool remove_edge(int from, int to) {
        int v0 = std::min(from, to), v1 = std::max(from, to);
        for (auto it = cns.begin(); it != cns.end(); ++it) {
            int t0 = std::min(it->from, it->to), t1 = std::max(it->from, it->to);
            if (v0 != t0 || v1 != t1) continue;
            cns.erase(it);
            return true;
        }
        return false;
    }
    // addEdgeToCnS, src/consensus.cpp:100-121 (stable sort by max id)
    void add_edge(int from, int to, const double* meas, const double* info) {
        int v0 = std::min(from, to), v1 = std::max(from, to);
        for (auto& t : cns) if (v0 == std::min(t.from, t.to) && v1 == std::max(t.from, t.to)) return;
        HostEdge e; e.from = from; e.to = to; e.meas.assign(meas, meas + mw); e.info.assign(info, info + d * d);
        cns.push_back(std::move(e));
        std::stable_sort(cns.begin(), cns.end(), [](const HostEdge& a, const HostEdge& b) { return std::max(a.from, a.to) < std::max(b.from, b.to); });
    }
    // computeIndependentSubgraph, src/consensus.cpp:123-171
    std::pair<int, int> independent_subgraph(int from, int to, std::vector<int>& members) const {
        int lo = std::min(from, to), hi = std::max(from, to);
        std::vector<char> inc(cns.size(), 0);
        bool found = true;
        while (found) {
            found = false;
            for (size_t k = 0; k < cns.size(); ++k) {
                if (inc[k]) continue;
                int t0 = std::min(cns[k].from, cns[k].to), t1 = std::max(cns[k].from, cns[k].to);
                if (std::min(t1, hi) - std::max(t0, lo) <= 0) continue;
                lo = std::min(lo, t0); hi = std::max(hi, t1);
                inc[k] = 1; found = true; members.push_back((int)k);
            }
        }
        return {lo, hi};
    }
};

}  // namespace ipcb
