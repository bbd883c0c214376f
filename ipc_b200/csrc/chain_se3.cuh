// chain_se3.cuh — SE(3) window check (EdgeSE3 / VertexSE3 instantiation of the reference, src/consensus.cpp:175), one CTA
// per check, K = 1 or 2 loop edges. Same structure as chain_se2.cuh (one sweep per Dogleg trial around a state machine),
// with g2o's SE(3) conventions:
//   estimate X = (R, t);  VertexSE3::oplusImpl: X <- X * [t = u(0:3), q = (sqrt(1 - |v|^2), v = u(3:6))]   (right / body increment)
//   EdgeSE3 error e = toVectorMQT(Z^-1 Xi^-1 Xj) = [t_E ; vec(q_E)], q_E normalised with w >= 0;  chi2 = e^T Omega e.
// Twist coordinates: a global twist Xi = (rho, phi) acts as X -> (I + Xi^) X, i.e. u = T Xi with
//   T = [R^T  -R^T [t]x ; 0  1/2 R^T].
// A common twist of both end points leaves the residual unchanged (Ji Ti + Jj Tj = 0), so the residual change of edge
// i -> j is A (Xi_j - Xi_i) with A = Jj Tj, Jj = blockdiag(R_E, w I + [v]x) (q_E = (w, v)), and
//   Q = A^-1 = [R_j R_E^T   [t_j]x N ; 0  N],   N = 2 R_j (w I + [v]x)^-1 = 2 R_j (w^2 I + v v^T - w [v]x) / w.
// Everything else (prefix sums PM = sum Q V Q^T, Pm = -sum Q e, the SPD force system, the model value, the Dogleg
// bookkeeping in g2o's vertex coordinates) is the 6-dimensional copy of chain_se2.cuh.
#pragma once
#include "chain_se2.cuh"

namespace ipcb {
namespace se3 {

constexpr int D6 = 6;
constexpr int NS6 = 21;              // packed symmetric 6x6, row-major upper triangle
constexpr int NP3 = NS6 + D6;        // prefix quantities: PM (21), Pm (6)
constexpr int SPECW3 = NP3 + 7;      // published per special vertex: prefix + pose (t, q)
constexpr int ODOM_REC3 = 7 + NS6 + NS6;   // zinv (t, q = w x y z), Omega (21), V = Omega^-1 (21)

struct LoopRec3 {
    int from, to;
    double zinv[7];
    double Om[NS6];
    double V[NS6];
};

IPC_HD constexpr int sidx(int r, int c) { return r <= c ? (r * 6 - r * (r - 1) / 2 + (c - r)) : (c * 6 - c * (c - 1) / 2 + (r - c)); }

struct P3 { double t[3]; double q[4]; };   // q = (w, x, y, z), unit

IPC_HD void q_mul(const double* a, const double* b, double* o) {
    const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    const double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
IPC_HD void q_to_R(const double* q, double* m) {   // Eigen toRotationMatrix, row-major
    const double tx = 2 * q[1], ty = 2 * q[2], tz = 2 * q[3];
    const double twx = tx * q[0], twy = ty * q[0], twz = tz * q[0];
    const double txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
    const double tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
    m[0] = 1 - (tyy + tzz); m[1] = txy - twz;       m[2] = txz + twy;
    m[3] = txy + twz;       m[4] = 1 - (txx + tzz); m[5] = tyz - twx;
    m[6] = txz - twy;       m[7] = tyz + twx;       m[8] = 1 - (txx + tyy);
}
IPC_HD void m3_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
IPC_HD void m3_mul_bt(const double* A, const double* B, double* C) {   // A * B^T
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c * 3] + A[r * 3 + 1] * B[c * 3 + 1] + A[r * 3 + 2] * B[c * 3 + 2];
}
IPC_HD void m3_vec(const double* A, const double* x, double* y) {
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = A[r * 3] * x[0] + A[r * 3 + 1] * x[1] + A[r * 3 + 2] * x[2];
}
IPC_HD void m3t_vec(const double* A, const double* x, double* y) {   // A^T x
#pragma unroll
    for (int r = 0; r < 3; ++r) y[r] = A[r] * x[0] + A[3 + r] * x[1] + A[6 + r] * x[2];
}
IPC_HD void cross3(const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
IPC_HD void skew_mul(const double* t, const double* A, double* C) {   // [t]x A
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double a0 = A[c], a1 = A[3 + c], a2 = A[6 + c];
        C[c] = t[1] * a2 - t[2] * a1; C[3 + c] = t[2] * a0 - t[0] * a2; C[6 + c] = t[0] * a1 - t[1] * a0;
    }
}
IPC_HD void sym6_vec(const double* S, const double* x, double* y) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = 0;
#pragma unroll
        for (int c = 0; c < 6; ++c) s += S[sidx(r, c)] * x[c];
        y[r] = s;
    }
}
IPC_HD double sym6_quad(const double* S, const double* x) { double y[6]; sym6_vec(S, x, y); return x[0] * y[0] + x[1] * y[1] + x[2] * y[2] + x[3] * y[3] + x[4] * y[4] + x[5] * y[5]; }

// a^-1 * b
IPC_HD void se3_rel(const P3& a, const P3& b, P3& o) {
    const double qc[4] = {a.q[0], -a.q[1], -a.q[2], -a.q[3]};
    q_mul(qc, b.q, o.q);
    double R[9]; q_to_R(a.q, R);
    const double d[3] = {b.t[0] - a.t[0], b.t[1] - a.t[1], b.t[2] - a.t[2]};
    m3t_vec(R, d, o.t);
}
// a * b
IPC_HD void se3_mul(const P3& a, const P3& b, P3& o) {
    double R[9]; q_to_R(a.q, R);
    double t[3]; m3_vec(R, b.t, t);
    double q[4]; q_mul(a.q, b.q, q);
    o.t[0] = t[0] + a.t[0]; o.t[1] = t[1] + a.t[1]; o.t[2] = t[2] + a.t[2];
    o.q[0] = q[0]; o.q[1] = q[1]; o.q[2] = q[2]; o.q[3] = q[3];
}
IPC_HD void q_normalize(double* q) {
    const double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
}

// linearisation of one edge i -> j
struct Lin3 {
    double e[6], we[6], chi;
    double RE[9], qe[4];        // error rotation, normalised quaternion with w >= 0
    double Rj[9], tj[3];        // to-vertex
    double sgn;                 // sign applied to q_zinv * q_B to make w >= 0
    P3 B;                       // Xi^-1 Xj
};
IPC_HD void lin3(const double* zinv, const P3& xi, const P3& xj, const double* Om, Lin3& L) {
    se3_rel(xi, xj, L.B);
    P3 Z; Z.t[0] = zinv[0]; Z.t[1] = zinv[1]; Z.t[2] = zinv[2]; Z.q[0] = zinv[3]; Z.q[1] = zinv[4]; Z.q[2] = zinv[5]; Z.q[3] = zinv[6];
    P3 E; se3_mul(Z, L.B, E);
    q_normalize(E.q);
    L.sgn = 1.0;
    if (E.q[0] < 0) { E.q[0] = -E.q[0]; E.q[1] = -E.q[1]; E.q[2] = -E.q[2]; E.q[3] = -E.q[3]; L.sgn = -1.0; }
#pragma unroll
    for (int q = 0; q < 4; ++q) L.qe[q] = E.q[q];
    L.e[0] = E.t[0]; L.e[1] = E.t[1]; L.e[2] = E.t[2]; L.e[3] = E.q[1]; L.e[4] = E.q[2]; L.e[5] = E.q[3];
    sym6_vec(Om, L.e, L.we);
    L.chi = 0;
#pragma unroll
    for (int q = 0; q < 6; ++q) L.chi += L.e[q] * L.we[q];
    q_to_R(E.q, L.RE);
    q_to_R(xj.q, L.Rj);
    L.tj[0] = xj.t[0]; L.tj[1] = xj.t[1]; L.tj[2] = xj.t[2];
}
// chi2 only
IPC_HD double chi3(const double* zinv, const P3& xi, const P3& xj, const double* Om) { Lin3 L; lin3(zinv, xi, xj, Om, L); return L.chi; }

// terms of one edge for the prefix sums: t[0..21) = Q V Q^T (packed), t[21..27) = -Q e
// If zf != nullptr also returns in *gain the edge's part of the predicted gain |J h_gn|^2_Omega: the residual change of an
// odometry edge under the force z of its region is -(e + V Q^T z); for a loop edge (sigma given) it is sigma V Q^T z - e.
IPC_HD void edge_terms3(const Lin3& L, const double* V, double* t, const double* zf = nullptr, const double* Om = nullptr, double sigma = 0.0,
                        double* gain = nullptr) {
    double Bp[9]; m3_mul_bt(L.Rj, L.RE, Bp);                     // R_j R_E^T
    const double w = L.qe[0], x = L.qe[1], y = L.qe[2], z = L.qe[3];
    const double iw = 1.0 / w;
    // (w I + [v]x)^-1 = (w^2 I + v v^T - w [v]x) / w
    const double Mi[9] = {(w * w + x * x) * iw, (x * y + w * z) * iw, (x * z - w * y) * iw,
                          (x * y - w * z) * iw, (w * w + y * y) * iw, (y * z + w * x) * iw,
                          (x * z + w * y) * iw, (y * z - w * x) * iw, (w * w + z * z) * iw};
    double N[9]; m3_mul(L.Rj, Mi, N);
#pragma unroll
    for (int q = 0; q < 9; ++q) N[q] *= 2.0;
    double S[9]; skew_mul(L.tj, N, S);                           // [t_j]x N
    // V blocks
    double Vtt[9], Vtr[9], Vrr[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) { Vtt[r * 3 + c] = V[sidx(r, c)]; Vtr[r * 3 + c] = V[sidx(r, 3 + c)]; Vrr[r * 3 + c] = V[sidx(3 + r, 3 + c)]; }
    double A1[9], A2[9], Tt[9], Tr[9], Br[9];
    m3_mul(Bp, Vtt, A1); m3_mul_bt(S, Vtr, A2);                  // S Vrt = S Vtr^T
#pragma unroll
    for (int q = 0; q < 9; ++q) Tt[q] = A1[q] + A2[q];           // (QV) top-left
    m3_mul(Bp, Vtr, A1); m3_mul(S, Vrr, A2);
#pragma unroll
    for (int q = 0; q < 9; ++q) Tr[q] = A1[q] + A2[q];           // (QV) top-right
    m3_mul(N, Vrr, Br);                                          // (QV) bottom-right
    double Mtt[9], Mtr[9], Mrr[9];
    m3_mul_bt(Tt, Bp, A1); m3_mul_bt(Tr, S, A2);
#pragma unroll
    for (int q = 0; q < 9; ++q) Mtt[q] = A1[q] + A2[q];
    m3_mul_bt(Tr, N, Mtr);
    m3_mul_bt(Br, N, Mrr);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c >= r) { t[sidx(r, c)] = Mtt[r * 3 + c]; t[sidx(3 + r, 3 + c)] = Mrr[r * 3 + c]; }
            t[sidx(r, 3 + c)] = Mtr[r * 3 + c];
        }
    double a[3], b[3], c3[3];
    m3_vec(Bp, L.e, a); m3_vec(S, L.e + 3, b); m3_vec(N, L.e + 3, c3);
    t[NS6 + 0] = -(a[0] + b[0]); t[NS6 + 1] = -(a[1] + b[1]); t[NS6 + 2] = -(a[2] + b[2]);
    t[NS6 + 3] = -c3[0]; t[NS6 + 4] = -c3[1]; t[NS6 + 5] = -c3[2];
    if (zf) {
        // yq = Q^T z = [Bp^T z_t ; S^T z_t + N^T z_r]
        double yq[6], y2[3];
        m3t_vec(Bp, zf, yq); m3t_vec(S, zf, yq + 3); m3t_vec(N, zf + 3, y2);
        yq[3] += y2[0]; yq[4] += y2[1]; yq[5] += y2[2];
        double vy[6]; sym6_vec(V, yq, vy);
        double w[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) w[q] = (sigma == 0.0) ? (L.e[q] + vy[q]) : (sigma * vy[q] - L.e[q]);
        *gain = sym6_quad(Om, w);
    }
}

// explicit Jacobians of the edge error w.r.t. the increments of Xi and Xj (steepest-descent path only)
IPC_HD void jac3(const double* zinv, const Lin3& L, double* Ji, double* Jj) {
#pragma unroll
    for (int q = 0; q < 36; ++q) { Ji[q] = 0; Jj[q] = 0; }
    const double w = L.qe[0], x = L.qe[1], y = L.qe[2], z = L.qe[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Jj[r * 6 + c] = L.RE[r * 3 + c];
    Jj[3 * 6 + 3] = w;  Jj[3 * 6 + 4] = -z; Jj[3 * 6 + 5] = y;
    Jj[4 * 6 + 3] = z;  Jj[4 * 6 + 4] = w;  Jj[4 * 6 + 5] = -x;
    Jj[5 * 6 + 3] = -y; Jj[5 * 6 + 4] = x;  Jj[5 * 6 + 5] = w;
    double Rz[9]; q_to_R(zinv + 3, Rz);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ji[r * 6 + c] = -Rz[r * 3 + c];
    const double* tb = L.B.t;
    const double tbx[9] = {0, -tb[2], tb[1], tb[2], 0, -tb[0], -tb[1], tb[0], 0};
    double Rt[9]; m3_mul(Rz, tbx, Rt);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ji[r * 6 + 3 + c] = 2 * Rt[r * 3 + c];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double dv[4] = {0, c == 0 ? -1.0 : 0.0, c == 1 ? -1.0 : 0.0, c == 2 ? -1.0 : 0.0};
        double t1[4], d[4];
        q_mul(zinv + 3, dv, t1); q_mul(t1, L.B.q, d);
        Ji[3 * 6 + 3 + c] = L.sgn * d[1]; Ji[4 * 6 + 3 + c] = L.sgn * d[2]; Ji[5 * 6 + 3 + c] = L.sgn * d[3];
    }
}
IPC_HD void m6t_vec(const double* J, const double* x, double* y) {   // J^T x
#pragma unroll
    for (int c = 0; c < 6; ++c) { double s = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s += J[r * 6 + c] * x[r];
        y[c] = s; }
}
IPC_HD void m6_vec(const double* J, const double* x, double* y) {
#pragma unroll
    for (int r = 0; r < 6; ++r) { double s = 0;
#pragma unroll
        for (int c = 0; c < 6; ++c) s += J[r * 6 + c] * x[c];
        y[r] = s; }
}

// VertexSE3::oplusImpl
IPC_HD void oplus3(const P3& p, const double* u, P3& o) {
    double R[9]; q_to_R(p.q, R);
    double dt[3]; m3_vec(R, u, dt);
    o.t[0] = p.t[0] + dt[0]; o.t[1] = p.t[1] + dt[1]; o.t[2] = p.t[2] + dt[2];
    const double w2 = 1.0 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
    if (w2 >= 0) { const double qu[4] = {sqrt(w2), u[3], u[4], u[5]}; q_mul(p.q, qu, o.q); }
    else { o.q[0] = p.q[0]; o.q[1] = p.q[1]; o.q[2] = p.q[2]; o.q[3] = p.q[3]; }
}

// ---- dense SPD solve (Cholesky, no pivoting), N <= 12, full row-major matrix, b overwritten -----------------------
template <int N> IPC_HD void chol_solve(double* A, double* b) {
    for (int j = 0; j < N; ++j) {
        double d = A[j * N + j];
        for (int k = 0; k < j; ++k) d -= A[j * N + k] * A[j * N + k];
        d = sqrt(d);
        A[j * N + j] = d;
        const double id = 1.0 / d;
        for (int i = j + 1; i < N; ++i) {
            double s = A[i * N + j];
            for (int k = 0; k < j; ++k) s -= A[i * N + k] * A[j * N + k];
            A[i * N + j] = s * id;
        }
    }
    for (int i = 0; i < N; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i * N + k] * b[k]; b[i] = s / A[i * N + i]; }
    for (int i = N - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < N; ++k) s -= A[k * N + i] * b[k]; b[i] = s / A[i * N + i]; }
}

// ---- collectives (27 scanned values) ----------------------------------------------------------------------------------
template <int NT, int NS> struct ScanSumMax3 {
    static constexpr int W = NP3 + NS + 1;
    static constexpr int NW = NT / 32 > 0 ? NT / 32 : 1;
    IPC_HD static void run(double* v, double* s, double& mx, double* red) {
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        double inc[NP3];
#pragma unroll
        for (int m = 0; m < NP3; ++m) {
            double x = v[m];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { double y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            inc[m] = x;
        }
#pragma unroll
        for (int m = 0; m < NS; ++m)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[m] += __shfl_xor_sync(0xffffffffu, s[m], o);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (NT <= 32) {
#pragma unroll
            for (int m = 0; m < NP3; ++m) v[m] = inc[m] - v[m];
            return;
        }
        if (lane == 31) {
#pragma unroll
            for (int m = 0; m < NP3; ++m) red[w * W + m] = inc[m];
#pragma unroll
            for (int m = 0; m < NS; ++m) red[w * W + NP3 + m] = s[m];
            red[w * W + NP3 + NS] = mx;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < NP3; ++m) {
            double base = 0;
            for (int i = 0; i < w; ++i) base += red[i * W + m];
            v[m] = base + inc[m] - v[m];
        }
#pragma unroll
        for (int m = 0; m < NS; ++m) { double t = 0; for (int i = 0; i < NW; ++i) t += red[i * W + NP3 + m]; s[m] = t; }
        double t = red[NP3 + NS];
        for (int i = 1; i < NW; ++i) t = fmax(t, red[i * W + NP3 + NS]);
        mx = t;
#else
        if (NT == 1) { for (int m = 0; m < NP3; ++m) v[m] = 0; return; }
        static_assert(NT * W <= 16 * (NP3 + 4), "host emulation: staging holds 16 threads");
        const int t = hd_tid();          // host emulation of the CTA (chain_se2.cuh, HostCta): sums in thread order
        for (int m = 0; m < NP3; ++m) red[t * W + m] = v[m];
        for (int m = 0; m < NS; ++m) red[t * W + NP3 + m] = s[m];
        red[t * W + NP3 + NS] = mx;
        bsync<NT>();
        for (int m = 0; m < NP3; ++m) { double b = 0; for (int i = 0; i < t; ++i) b += red[i * W + m]; v[m] = b; }
        for (int m = 0; m < NS; ++m) { double b = 0; for (int i = 0; i < NT; ++i) b += red[i * W + NP3 + m]; s[m] = b; }
        double b = red[NP3 + NS];
        for (int i = 1; i < NT; ++i) b = fmax(b, red[i * W + NP3 + NS]);
        mx = b;
#endif
    }
};

// ---- per-check working set ------------------------------------------------------------------------------------------
struct StepSpec3 { double z[3][6]; double C[3][6]; double model; double gain_loops; int rs, re; };
struct CheckGeom3 { int K, lo, L, rs, re, first_is_c, last_is_c, c_a_is_rs, c_b_is_L, m_a_is_rs, m_b_is_L; };
struct UniBlock3 { LoopRec3 Lc, Lm; StepSpec3 sol; double n_c, n_m; CheckGeom3 g;
                   double lt[2][32];        // per loop (candidate, member), staged by lanes 0 / 1 of warp 0: prefix terms t (27), chi2, sigma, gain part
                   double zs[12]; };        // forces (zc, zm) of the last solve
constexpr int RED3_DOUBLES = 16 * (NP3 + 4);
constexpr int UNI3_DOUBLES = (sizeof(UniBlock3) + 7) / 8;
constexpr int CHAIN3_SMALL_DOUBLES = 2 * RED3_DOUBLES + 2 * NSPEC * SPECW3 + UNI3_DOUBLES + 16 * 7 + 32 * 7;   // + dead-reckoning staging
constexpr int CHAIN3_STATE = 7;      // per-vertex doubles of state: t (3), q (4)
constexpr int CHAIN3_SCRATCH = 7 + 12;   // per-vertex doubles of global scratch: backup (7), b (6), h_gn (6)
IPC_HD constexpr int global_state3_doubles(int capv, int nt) { return 7 * (capv + 4 * nt); }   // step tiles: (S + 1) * 7 * NT <= 7 (L + 3 NT)

// State record of a vertex (t, q: 7 doubles). Shared-memory state (gst = 0): AoS by vertex, components 1 double apart. Global-memory state
// (gst = 1: the one-warp-per-check kernels): tiles of 7 x nt doubles by STEP as in chain_se2.cuh (StateAt) — the vertex thread t reaches at
// step i of its segment walk is element t of tile i + 1, components nt doubles apart, tile 0 holds the fixed origin — so every access of a
// warp is one contiguous run; a trial sweep writes its new poses to the second buffer stw (accepted: swap; rejected: nothing to undo).
struct ChainMem3 {
    double* st; double* scr; double* small; int capv;
    double* stw;             // == st for shared-memory state
    int gst, nt, S, cs;      // layout: global tiles?, threads, segment length (vertices per thread), component stride
    IPC_HD size_t tile_off(int j) const { if (j == 0) return 0; const int t = (j - 1) / S, i = (j - 1) - t * S; return ((size_t)(i + 1) * 7) * nt + t; }
    IPC_HD double* P(int j) const { return gst ? st + tile_off(j) : st + 7 * j; }
    IPC_HD double* PW(int j) const { return gst ? stw + tile_off(j) : stw + 7 * j; }
    IPC_HD double* Pit(int j, int i, int t) const { return gst ? st + ((size_t)(i + 1) * 7) * nt + t : st + 7 * j; }        // vertex j = k0(t) + 1 + i
    IPC_HD double* PWit(int j, int i, int t) const { return gst ? stw + ((size_t)(i + 1) * 7) * nt + t : stw + 7 * j; }
    IPC_HD double* B(int j) const { return scr + 7 * j; }
    IPC_HD double* G(int j) const { return scr + 7 * (size_t)capv + 12 * j; }
    IPC_HD double* red() const { return small; }
    IPC_HD double* spec() const { return small + 2 * RED3_DOUBLES; }
    IPC_HD UniBlock3* U() const { return reinterpret_cast<UniBlock3*>(small + 2 * RED3_DOUBLES + 2 * NSPEC * SPECW3); }
    IPC_HD double* stage() const { return small + 2 * RED3_DOUBLES + 2 * NSPEC * SPECW3 + UNI3_DOUBLES; }
};
IPC_HD void load_pose(const double* p, P3& o, int cs = 1) { o.t[0] = p[0]; o.t[1] = p[cs]; o.t[2] = p[2 * cs]; o.q[0] = p[3 * cs]; o.q[1] = p[4 * cs]; o.q[2] = p[5 * cs]; o.q[3] = p[6 * cs]; }
IPC_HD void store_pose(double* p, const P3& o, int cs = 1) { p[0] = o.t[0]; p[cs] = o.t[1]; p[2 * cs] = o.t[2]; p[3 * cs] = o.q[0]; p[4 * cs] = o.q[1]; p[5 * cs] = o.q[2]; p[6 * cs] = o.q[3]; }

struct ThreadState3 { int k0, k1; P3 pa; double base[NP3]; };

IPC_HD void twist_at3(const StepSpec3* sp, int j, const double* pre, double* Xi) {
    const int r = (j <= sp->rs) ? 0 : (j <= sp->re ? 1 : 2);
    double y[6]; sym6_vec(pre, sp->z[r], y);
#pragma unroll
    for (int q = 0; q < 6; ++q) Xi[q] = pre[NS6 + q] - y[q] - sp->C[r][q];
}
// u = T Xi at the vertex with pose p: u_t = R^T (rho - t x phi), u_r = 1/2 R^T phi
IPC_HD void gn_step_at3(const StepSpec3* sp, int j, const double* pre, const P3& p, double* u) {
    double Xi[6]; twist_at3(sp, j, pre, Xi);
    double R[9]; q_to_R(p.q, R);
    double c[3]; cross3(p.t, Xi + 3, c);
    const double a[3] = {Xi[0] - c[0], Xi[1] - c[1], Xi[2] - c[2]};
    m3t_vec(R, a, u);
    double r[3]; m3t_vec(R, Xi + 3, r);
    u[3] = 0.5 * r[0]; u[4] = 0.5 * r[1]; u[5] = 0.5 * r[2];
}

struct SweepOut3 { double chi, mx, hh, gain; };

// Two loops per sweep: loop A applies the step (old linearisation -> running prefix -> u -> new pose, backup), loop B
// re-linearises at the new poses read back from shared memory. With 27 prefix values and a 46-double linearisation per
// edge, keeping both running prefixes alive in one loop spills hundreds of bytes per thread even at 255 registers.
template <int NT> IPC_HD void sweep3(const ChainMem3& M, const double* odom, int mode, double c1, double c2, bool acc_gain, ThreadState3& ts, SweepOut3& out,
                                     int& buf, const int* spec_v) {
    const int k0 = ts.k0, k1 = ts.k1, tid_ = hd_tid();
    const StepSpec3* sp = &M.U()->sol;
    double* spec = M.spec() + (size_t)buf * NSPEC * SPECW3;
    double hh = 0, gain = 0;
    if (mode != STEP_NONE) {
        double pre[NP3];
#pragma unroll
        for (int m = 0; m < NP3; ++m) pre[m] = ts.base[m];
        P3 oa = ts.pa;
        if (k0 > 0 && k0 < k1) {     // boundary vertex k0: same arithmetic as its owner => identical bits
            double u[6];
            if (mode == STEP_GN) gn_step_at3(sp, k0, pre, oa, u);
            else { const double* gq = M.G(k0);
#pragma unroll
                for (int q = 0; q < 6; ++q) u[q] = c1 * gq[q] + c2 * gq[6 + q]; }
            oplus3(oa, u, ts.pa);
        }
        for (int k = k0; k < k1; ++k) {
            const int j = k + 1;
            const double* rec = odom + (size_t)ODOM_REC3 * k;
            P3 ob; load_pose(M.Pit(j, k - k0, tid_), ob, M.cs);
            double u[6];
            if (mode == STEP_GN) {
                Lin3 eo; lin3(rec, oa, ob, rec + 7, eo);
                const int rg = (k < sp->rs) ? 0 : (k < sp->re ? 1 : 2);
                double to[NP3], ge = 0; edge_terms3(eo, rec + 7 + NS6, to, acc_gain ? sp->z[rg] : nullptr, rec + 7, 0.0, &ge);
                gain += ge;
#pragma unroll
                for (int m = 0; m < NP3; ++m) pre[m] += to[m];
                gn_step_at3(sp, j, pre, ob, u);
            } else { const double* gq = M.G(j);
#pragma unroll
                for (int q = 0; q < 6; ++q) u[q] = c1 * gq[q] + c2 * gq[6 + q]; }
            if (!M.gst) store_pose(M.B(j), ob);
            P3 nb; oplus3(ob, u, nb);
#pragma unroll
            for (int q = 0; q < 6; ++q) hh += u[q] * u[q];
            store_pose(M.PWit(j, k - k0, tid_), nb, M.cs);
            oa = ob;
        }
    }
    double run[NP3];
#pragma unroll
    for (int m = 0; m < NP3; ++m) run[m] = 0;
    double chi = 0, mx = 0;
    bool has_spec = false;
#pragma unroll
    for (int q = 1; q < NSPEC; ++q) has_spec |= (spec_v[q] > k0 && spec_v[q] <= k1);
    {
        P3 na = ts.pa;
        for (int k = k0; k < k1; ++k) {
            const int j = k + 1;
            const double* rec = odom + (size_t)ODOM_REC3 * k;
            P3 nb; load_pose(mode != STEP_NONE ? M.PWit(j, k - k0, tid_) : M.Pit(j, k - k0, tid_), nb, M.cs);
            Lin3 e; lin3(rec, na, nb, rec + 7, e);
            double t[NP3]; edge_terms3(e, rec + 7 + NS6, t);
            chi += e.chi; mx = fmax(mx, e.chi);
#pragma unroll
            for (int m = 0; m < NP3; ++m) run[m] += t[m];
            if (has_spec) {
#pragma unroll
                for (int q = 1; q < NSPEC; ++q)
                    if (j == spec_v[q]) { double* o = spec + q * SPECW3;
#pragma unroll
                        for (int m = 0; m < NP3; ++m) o[m] = run[m];
                        store_pose(o + NP3, nb); }
            }
            na = nb;
        }
    }
    double s[3] = {chi, hh, gain};
    ScanSumMax3<NT, 3>::run(run, s, mx, M.red() + (size_t)buf * RED3_DOUBLES);
#pragma unroll
    for (int m = 0; m < NP3; ++m) ts.base[m] = run[m];
    out.chi = s[0]; out.hh = s[1]; out.gain = s[2]; out.mx = mx;
    if (has_spec) {
#pragma unroll
        for (int q = 1; q < NSPEC; ++q) { const int v = spec_v[q];
            if (v > k0 && v <= k1) { double* o = spec + q * SPECW3;
#pragma unroll
                for (int m = 0; m < NP3; ++m) o[m] += run[m]; } }
    }
    bsync<NT>();
    buf ^= 1;
}

template <int NT> IPC_HD void rollback3(const ChainMem3& M, ThreadState3& ts) {
    if (!M.gst) for (int k = ts.k0; k < ts.k1; ++k) { const int j = k + 1; const double* b = M.B(j); double* p = M.P(j);
#pragma unroll
        for (int q = 0; q < 7; ++q) p[q] = b[q]; }
    bsync<NT>();
    if (ts.k0 < ts.k1 && ts.k0 > 0) load_pose(M.P(ts.k0), ts.pa, M.cs);
    bsync<NT>();
}

struct LoopNow3 { Lin3 e; double t[NP3]; double sigma; };

// thread 0: loops at the published state, and (if kept) the GN solution of the new linearisation
IPC_HD_COLD void eval_and_solve3_t0(ChainMem3 M, int buf, double odom_chi, double cur_chi, double linearGain, bool force) {
    UniBlock3* U = M.U();
    const CheckGeom3& g = U->g;
    const double* spec = M.spec() + (size_t)(buf ^ 1) * NSPEC * SPECW3;
    const double* o1 = spec + 1 * SPECW3; const double* o2 = spec + 2 * SPECW3; const double* o3 = spec + 3 * SPECW3;
    const bool z1 = g.rs == 0;
    double pre1[NP3], pre2[NP3], pre3[NP3];
    for (int m = 0; m < NP3; ++m) { pre1[m] = z1 ? 0.0 : o1[m]; pre2[m] = o2[m]; pre3[m] = o3[m]; }
    P3 org; org.t[0] = org.t[1] = org.t[2] = 0; org.q[0] = 1; org.q[1] = org.q[2] = org.q[3] = 0;
    P3 p1 = org, p2, p3;
    if (!z1) load_pose(o1 + NP3, p1);
    load_pose(o2 + NP3, p2); load_pose(o3 + NP3, p3);
    LoopNow3 lc, lm;
    {
        const P3& pa = g.c_a_is_rs ? p1 : org; const P3& pb = g.c_b_is_L ? p3 : p2;
        const bool to_hi = U->Lc.to > U->Lc.from;
        lin3(U->Lc.zinv, to_hi ? pa : pb, to_hi ? pb : pa, U->Lc.Om, lc.e);
        edge_terms3(lc.e, U->Lc.V, lc.t); lc.sigma = to_hi ? 1.0 : -1.0;
    }
    double c = lc.e.chi, m = 0;
    if (g.K == 2) {
        const P3& pa = g.m_a_is_rs ? p1 : org; const P3& pb = g.m_b_is_L ? p3 : p2;
        const bool to_hi = U->Lm.to > U->Lm.from;
        lin3(U->Lm.zinv, to_hi ? pa : pb, to_hi ? pb : pa, U->Lm.Om, lm.e);
        edge_terms3(lm.e, U->Lm.V, lm.t); lm.sigma = to_hi ? 1.0 : -1.0;
        m = lm.e.chi;
    }
    if (fabs(linearGain) < 1e-12) linearGain = 1e-12;
    const double rho = (cur_chi - (odom_chi + c + m)) / linearGain;
    U->n_c = c; U->n_m = m;
    if (!(force || rho > 0)) return;
    // ---- GN solve: (P_ll' + delta W_l) z_l' = q_l + sigma_l Q_l e_l
    StepSpec3* sp = &U->sol;
    double acc[3][NP3];
    for (int q = 0; q < NP3; ++q) { acc[0][q] = pre1[q]; acc[1][q] = pre2[q] - pre1[q]; acc[2][q] = pre3[q] - pre2[q]; }
    double zc[6], zm[6] = {0, 0, 0, 0, 0, 0};
    if (g.K == 1) {
        double A[36], r[6];
        for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) A[i * 6 + j] = acc[1][sidx(i, j)] + lc.t[sidx(i, j)]; r[i] = acc[1][NS6 + i] - lc.sigma * lc.t[NS6 + i]; }
        chol_solve<6>(A, r);
        for (int i = 0; i < 6; ++i) zc[i] = r[i];
    } else {
        double A[144], r[12];
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) {
                const int s = sidx(i, j);
                const double pcm = acc[1][s];
                const double pcc = acc[1][s] + (g.first_is_c ? acc[0][s] : 0.0) + (g.last_is_c ? acc[2][s] : 0.0) + lc.t[s];
                const double pmm = acc[1][s] + (g.first_is_c ? 0.0 : acc[0][s]) + (g.last_is_c ? 0.0 : acc[2][s]) + lm.t[s];
                A[i * 12 + j] = pcc; A[i * 12 + 6 + j] = pcm; A[(6 + i) * 12 + j] = pcm; A[(6 + i) * 12 + 6 + j] = pmm;
            }
            r[i] = acc[1][NS6 + i] + (g.first_is_c ? acc[0][NS6 + i] : 0.0) + (g.last_is_c ? acc[2][NS6 + i] : 0.0) - lc.sigma * lc.t[NS6 + i];
            r[6 + i] = acc[1][NS6 + i] + (g.first_is_c ? 0.0 : acc[0][NS6 + i]) + (g.last_is_c ? 0.0 : acc[2][NS6 + i]) - lm.sigma * lm.t[NS6 + i];
        }
        chol_solve<12>(A, r);
        for (int i = 0; i < 6; ++i) { zc[i] = r[i]; zm[i] = r[6 + i]; }
    }
    sp->rs = g.rs; sp->re = g.re;
    double z[3][6], C[3][6];
    for (int q = 0; q < 6; ++q) {
        if (g.K == 1) { z[0][q] = 0; z[1][q] = zc[q]; z[2][q] = 0; }
        else { z[0][q] = g.first_is_c ? zc[q] : zm[q]; z[1][q] = zc[q] + zm[q]; z[2][q] = g.last_is_c ? zc[q] : zm[q]; }
    }
    {
        double dz[6], t[6];
        for (int q = 0; q < 6; ++q) { C[0][q] = 0; dz[q] = z[0][q] - z[1][q]; }
        sym6_vec(pre1, dz, t);
        for (int q = 0; q < 6; ++q) { C[1][q] = t[q]; dz[q] = z[1][q] - z[2][q]; }
        sym6_vec(pre2, dz, t);
        for (int q = 0; q < 6; ++q) C[2][q] = C[1][q] + t[q];
    }
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 6; ++q) { sp->z[r][q] = z[r][q]; sp->C[r][q] = C[r][q]; }
    double model = sym6_quad(lc.t, zc);
    if (g.K == 2) model += sym6_quad(lm.t, zm);
    for (int r = 0; r < 3; ++r) model += sym6_quad(acc[r], z[r]);
    sp->model = model;
    {
        double tt[NP3], gl = 0, g2 = 0;
        edge_terms3(lc.e, U->Lc.V, tt, zc, U->Lc.Om, lc.sigma, &gl);
        if (g.K == 2) edge_terms3(lm.e, U->Lm.V, tt, zm, U->Lm.Om, lm.sigma, &g2);
        sp->gain_loops = gl + g2;
    }
}
#ifdef __CUDA_ARCH__
// Device form of eval_and_solve3_t0, executed by the 32 lanes of warp 0 (the CTA-per-check kernels spent 55 % of their warp samples at the
// barrier behind the one-thread version, ncu profiles/r02_ncu_se3.txt): lanes 0 / 1 linearise the candidate / member loop, lane r owns row r
// of the augmented 6K x (6K + 1) force system (assembled from the special-vertex table in shared memory) and the SPD system is solved by
// Gaussian elimination with the pivot row broadcast by shuffles (no pivoting needed: SPD), then lanes 0..5 form z / C component-wise.
template <int N> __device__ __forceinline__ double warp_spd_solve(double (&a)[13], int lane) {
    // elimination: after step p, rows r > p have a zero in column p
#pragma unroll
    for (int p = 0; p < N; ++p) {
        const double piv = __shfl_sync(0xffffffffu, a[p], p);
        const double f = (lane > p && lane < N) ? a[p] / piv : 0.0;
#pragma unroll
        for (int c = p + 1; c <= N; ++c) {
            const double pc = __shfl_sync(0xffffffffu, a[c == N ? 12 : c], p);
            a[c == N ? 12 : c] = fma(-f, pc, a[c == N ? 12 : c]);
        }
    }
    // back substitution: lane p produces x_p once x_{p+1..N-1} are known
    double x = 0.0;
#pragma unroll
    for (int p = N - 1; p >= 0; --p) {
        const double xp = __shfl_sync(0xffffffffu, a[12] / a[p], p);      // lane p: (rhs - sum_{c>p} a[c] x_c) / a[p], the sum already folded into a[12]
        if (lane == p) x = xp;
        if (lane < p) a[12] = fma(-a[p], xp, a[12]);
    }
    return x;
}
__device__ __noinline__ void eval_and_solve3_w0(ChainMem3 M, int buf, double odom_chi, double cur_chi, double linearGain, bool force) {
    UniBlock3* U = M.U();
    const CheckGeom3& g = U->g;
    const int lane = threadIdx.x & 31;
    const double* spec = M.spec() + (size_t)(buf ^ 1) * NSPEC * SPECW3;
    const double* o1 = spec + 1 * SPECW3; const double* o2 = spec + 2 * SPECW3; const double* o3 = spec + 3 * SPECW3;
    const bool z1 = g.rs == 0;
    // ---- lanes 0 / 1: the loop edges at the published state
    Lin3 le;
    const LoopRec3& Lp = lane == 0 ? U->Lc : U->Lm;
    double sigma = 0;
    if (lane < g.K) {
        P3 org; org.t[0] = org.t[1] = org.t[2] = 0; org.q[0] = 1; org.q[1] = org.q[2] = org.q[3] = 0;
        P3 p1 = org, p2, p3;
        if (!z1) load_pose(o1 + NP3, p1);
        load_pose(o2 + NP3, p2); load_pose(o3 + NP3, p3);
        const bool a_is_rs = lane == 0 ? g.c_a_is_rs : g.m_a_is_rs, b_is_L = lane == 0 ? g.c_b_is_L : g.m_b_is_L;
        const P3& pa = a_is_rs ? p1 : org; const P3& pb = b_is_L ? p3 : p2;
        const bool to_hi = Lp.to > Lp.from;
        lin3(Lp.zinv, to_hi ? pa : pb, to_hi ? pb : pa, Lp.Om, le);
        double t[NP3]; edge_terms3(le, Lp.V, t);
        sigma = to_hi ? 1.0 : -1.0;
        double* o = U->lt[lane];
#pragma unroll
        for (int q = 0; q < NP3; ++q) o[q] = t[q];
        o[27] = le.chi; o[28] = sigma;
    }
    __syncwarp();
    const double* tc = U->lt[0]; const double* tm = U->lt[1];
    const double c = tc[27], m = g.K == 2 ? tm[27] : 0.0;
    if (fabs(linearGain) < 1e-12) linearGain = 1e-12;
    const double rho = (cur_chi - (odom_chi + c + m)) / linearGain;
    if (lane == 0) { U->n_c = c; U->n_m = m; }
    if (!(force || rho > 0)) return;
    // ---- GN solve: (P_ll' + delta W_l) z_l' = q_l + sigma_l Q_l e_l; lane r assembles row r (region sums read on the fly:
    //      acc0 = pre1 (0 when rs == 0), acc1 = pre2 - pre1, acc2 = pre3 - pre2)
    StepSpec3* sp = &U->sol;
    auto A0 = [&](int q) { return z1 ? 0.0 : o1[q]; };
    auto A1 = [&](int q) { return o2[q] - (z1 ? 0.0 : o1[q]); };
    auto A2 = [&](int q) { return o3[q] - o2[q]; };
    double a[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) a[q] = 0.0;
    const int i6 = lane % 6, blk = lane / 6;
    double x;
    if (g.K == 1) {
        if (lane < 6) {
#pragma unroll
            for (int j = 0; j < 6; ++j) { const int sx = sidx(i6, j); a[j] = A1(sx) + tc[sx]; }
            a[12] = A1(NS6 + i6) - tc[28] * tc[NS6 + i6];
        }
        x = warp_spd_solve<6>(a, lane);
    } else {
        if (lane < 12) {
            const bool fc = g.first_is_c != 0, lc_ = g.last_is_c != 0;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int sx = sidx(i6, j);
                const double a0 = A0(sx), a1 = A1(sx), a2 = A2(sx);
                const double pcc = a1 + (fc ? a0 : 0.0) + (lc_ ? a2 : 0.0) + tc[sx];
                const double pmm = a1 + (fc ? 0.0 : a0) + (lc_ ? 0.0 : a2) + tm[sx];
                a[j] = blk == 0 ? pcc : a1;
                a[6 + j] = blk == 0 ? a1 : pmm;
            }
            const double r0 = A0(NS6 + i6), r1 = A1(NS6 + i6), r2 = A2(NS6 + i6);
            a[12] = blk == 0 ? r1 + (fc ? r0 : 0.0) + (lc_ ? r2 : 0.0) - tc[28] * tc[NS6 + i6]
                             : r1 + (fc ? 0.0 : r0) + (lc_ ? 0.0 : r2) - tm[28] * tm[NS6 + i6];
        }
        x = warp_spd_solve<12>(a, lane);
    }
    if (lane < 12) U->zs[lane] = (lane < 6 * g.K) ? x : 0.0;
    __syncwarp();
    const double* zc = U->zs; const double* zm = U->zs + 6;
    if (lane == 0) { sp->rs = g.rs; sp->re = g.re; }
    if (lane < 6) {     // component `lane` of z / C of the three regions
        const int q = lane;
        auto zreg = [&](int r, int p) { return g.K == 1 ? (r == 1 ? zc[p] : 0.0) : (r == 0 ? (g.first_is_c ? zc[p] : zm[p]) : (r == 1 ? zc[p] + zm[p] : (g.last_is_c ? zc[p] : zm[p]))); };
        double c1 = 0, c2 = 0;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const int sx = sidx(q, p);
            c1 += A0(sx) * (zreg(0, p) - zreg(1, p));                 // C_1 = PM(rs) (z_0 - z_1)
            c2 += o2[sx] * (zreg(1, p) - zreg(2, p));                 // C_2 = C_1 + PM(re) (z_1 - z_2)
        }
        sp->z[0][q] = zreg(0, q); sp->z[1][q] = zreg(1, q); sp->z[2][q] = zreg(2, q);
        sp->C[0][q] = 0; sp->C[1][q] = c1; sp->C[2][q] = c1 + c2;
    }
    __syncwarp();
    // model value: lanes 0..4 take one quadratic form each (loop c, loop m, regions 0..2), summed by shuffles
    double mq = 0;
    if (lane < 5) {
        double zz[6];
#pragma unroll
        for (int p = 0; p < 6; ++p) zz[p] = lane == 0 ? zc[p] : (lane == 1 ? zm[p] : sp->z[lane - 2][p]);
        if (!(lane == 1 && g.K == 1)) {
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double y = 0;
#pragma unroll
                for (int p = 0; p < 6; ++p) { const int sx = sidx(r, p); y += (lane == 0 ? tc[sx] : lane == 1 ? tm[sx] : lane == 2 ? A0(sx) : lane == 3 ? A1(sx) : A2(sx)) * zz[p]; }
                mq += zz[r] * y;
            }
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mq += __shfl_down_sync(0xffffffffu, mq, o, 8);
    // loop part of the predicted gain: lanes 0 / 1 again (their linearisation is still in registers)
    double gl = 0;
    if (lane < g.K) { double tt[NP3]; edge_terms3(le, Lp.V, tt, lane == 0 ? zc : zm, Lp.Om, sigma, &gl); }
    gl += __shfl_down_sync(0xffffffffu, gl, 1);
    if (lane == 0) { sp->model = mq; sp->gain_loops = gl; }
}
#endif
template <int NT> IPC_HD void eval_and_solve3(const ChainMem3& M, int buf, double odom_chi, double cur_chi, double linearGain, bool force, double& n_c, double& n_m) {
#ifdef __CUDA_ARCH__
    if (threadIdx.x < 32) eval_and_solve3_w0(M, buf, odom_chi, cur_chi, linearGain, force);
#else
    if (hd_tid() == 0) eval_and_solve3_t0(M, buf, odom_chi, cur_chi, linearGain, force);
#endif
    bsync<NT>();
    n_c = M.U()->n_c; n_m = M.U()->n_m;
}

template <int NT> IPC_HD double gn_norm_sq3(const ChainMem3& M, const double* odom, const ThreadState3& ts) {
    double v[1] = {0};
    const StepSpec3* sp = &M.U()->sol;
    double pre[NP3];
#pragma unroll
    for (int m = 0; m < NP3; ++m) pre[m] = ts.base[m];
    P3 pa = ts.pa;
    for (int k = ts.k0; k < ts.k1; ++k) {
        const int j = k + 1;
        const double* rec = odom + (size_t)ODOM_REC3 * k;
        P3 pb; load_pose(M.Pit(j, k - ts.k0, hd_tid()), pb, M.cs);
        Lin3 e; lin3(rec, pa, pb, rec + 7, e);
        double t[NP3]; edge_terms3(e, rec + 7 + NS6, t);
#pragma unroll
        for (int m = 0; m < NP3; ++m) pre[m] += t[m];
        double u[6]; gn_step_at3(sp, j, pre, pb, u);
#pragma unroll
        for (int q = 0; q < 6; ++q) v[0] += u[q] * u[q];
        pa = pb;
    }
    hd_block_sum<NT, 1>(v, M.stage());
    return v[0];
}

// steepest-descent sweeps: b_j and h_gn,j (g2o vertex coordinates) into the scratch; bb, bh, hh, bHb
template <int NT> IPC_HD void sd_sweeps3(const ChainMem3& M, const double* odom, const ThreadState3& ts, double& bb, double& bh, double& hh, double& bHb) {
    const CheckGeom3& g = M.U()->g;
    const int k0 = ts.k0, k1 = ts.k1, L = g.L;
    const StepSpec3* sp = &M.U()->sol;
    const LoopRec3& Lc = M.U()->Lc; const LoopRec3& Lm = M.U()->Lm;
    const int cjf = Lc.from - g.lo, cjt = Lc.to - g.lo; int mjf = -1, mjt = -1;
    double gci[6], gcj[6], gmi[6] = {0, 0, 0, 0, 0, 0}, gmj[6] = {0, 0, 0, 0, 0, 0};
    double Jci[36], Jcj[36], Jmi[36], Jmj[36];
    {
        P3 pf, pt; load_pose(M.P(cjf), pf, M.cs); load_pose(M.P(cjt), pt, M.cs);
        Lin3 e; lin3(Lc.zinv, pf, pt, Lc.Om, e); jac3(Lc.zinv, e, Jci, Jcj);
        m6t_vec(Jci, e.we, gci); m6t_vec(Jcj, e.we, gcj);
    }
    if (g.K == 2) {
        mjf = Lm.from - g.lo; mjt = Lm.to - g.lo;
        P3 pf, pt; load_pose(M.P(mjf), pf, M.cs); load_pose(M.P(mjt), pt, M.cs);
        Lin3 e; lin3(Lm.zinv, pf, pt, Lm.Om, e); jac3(Lm.zinv, e, Jmi, Jmj);
        m6t_vec(Jmi, e.we, gmi); m6t_vec(Jmj, e.we, gmj);
    }
    double v[3] = {0, 0, 0};
    double pre[NP3];
#pragma unroll
    for (int m = 0; m < NP3; ++m) pre[m] = ts.base[m];
    auto finish_vertex = [&](int j, const double* gsum, const P3& p) {
        double b[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) b[q] = -gsum[q];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            if (j == cjf) b[q] -= gci[q];
            if (j == cjt) b[q] -= gcj[q];
            if (j == mjf) b[q] -= gmi[q];
            if (j == mjt) b[q] -= gmj[q];
        }
        double u[6]; gn_step_at3(sp, j, pre, p, u);
        double* gq = M.G(j);
#pragma unroll
        for (int q = 0; q < 6; ++q) { gq[q] = b[q]; gq[6 + q] = u[q]; v[0] += b[q] * b[q]; v[1] += b[q] * u[q]; v[2] += u[q] * u[q]; }
    };
    if (k0 < k1) {
        P3 pa = ts.pa;
        double gprev[6] = {0, 0, 0, 0, 0, 0};
        for (int k = k0; k <= k1 && k < L; ++k) {
            const double* rec = odom + (size_t)ODOM_REC3 * k;
            P3 pb; load_pose(k < k1 ? M.Pit(k + 1, k - k0, hd_tid()) : M.P(k + 1), pb, M.cs);
            Lin3 e; lin3(rec, pa, pb, rec + 7, e);
            double Ji[36], Jj[36]; jac3(rec, e, Ji, Jj);
            double gi[6], gj[6]; m6t_vec(Ji, e.we, gi); m6t_vec(Jj, e.we, gj);
            if (k > k0) {
                double gs[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) gs[q] = gprev[q] + gi[q];
                finish_vertex(k, gs, pa);
            }
            double t[NP3]; edge_terms3(e, rec + 7 + NS6, t);
#pragma unroll
            for (int m = 0; m < NP3; ++m) pre[m] += t[m];
#pragma unroll
            for (int q = 0; q < 6; ++q) gprev[q] = gj[q];
            pa = pb;
        }
        if (k1 == L) finish_vertex(L, gprev, pa);
    }
    if (hd_tid() == 0) { double* g0 = M.G(0);
#pragma unroll
        for (int q = 0; q < 12; ++q) g0[q] = 0; }
    hd_block_sum<NT, 3>(v, M.stage());
    bsync<NT>();
    bb = v[0]; bh = v[1]; hh = v[2];
    double w[1] = {0};
    if (k0 < k1) {
        P3 pa = ts.pa;
        double ba[6];
        { const double* gq = M.G(k0);
#pragma unroll
          for (int q = 0; q < 6; ++q) ba[q] = gq[q]; }
        for (int k = k0; k < k1; ++k) {
            const double* rec = odom + (size_t)ODOM_REC3 * k;
            P3 pb; load_pose(M.Pit(k + 1, k - k0, hd_tid()), pb, M.cs);
            Lin3 e; lin3(rec, pa, pb, rec + 7, e);
            double Ji[36], Jj[36]; jac3(rec, e, Ji, Jj);
            double bv[6]; { const double* gq = M.G(k + 1);
#pragma unroll
                for (int q = 0; q < 6; ++q) bv[q] = gq[q]; }
            double q1[6], q2[6]; m6_vec(Ji, ba, q1); m6_vec(Jj, bv, q2);
#pragma unroll
            for (int q = 0; q < 6; ++q) q1[q] += q2[q];
            w[0] += sym6_quad(rec + 7, q1);
#pragma unroll
            for (int q = 0; q < 6; ++q) ba[q] = bv[q];
            pa = pb;
        }
    }
    if (hd_tid() == 0) {
        double q1[6], q2[6];
        m6_vec(Jci, M.G(cjf), q1); m6_vec(Jcj, M.G(cjt), q2);
        for (int q = 0; q < 6; ++q) q1[q] += q2[q];
        w[0] += sym6_quad(Lc.Om, q1);
        if (g.K == 2) {
            m6_vec(Jmi, M.G(mjf), q1); m6_vec(Jmj, M.G(mjt), q2);
            for (int q = 0; q < 6; ++q) q1[q] += q2[q];
            w[0] += sym6_quad(Lm.Om, q1);
        }
    }
    hd_block_sum<NT, 1>(w, M.stage());
    bHb = w[0];
}

// exclusive "scan" of rigid transforms over the threads (dead-reckoning): T_excl(t) = T_0 * T_1 * ... * T_{t-1}
template <int NT> IPC_HD void se3_excl_scan(const ChainMem3& M, const P3& mine, P3& excl) {
#ifdef __CUDA_ARCH__
    double* stage = M.stage();            // [16 warps][7] warp totals
    double* lane_tot = M.st;              // the pose array is still unused here and holds >= NT poses (host guarantees capv >= NT)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    store_pose(lane_tot + 7 * tid, mine);
    __syncthreads();
    if (lane == 0) {
        P3 acc; load_pose(lane_tot + 7 * tid, acc);
        for (int i = 1; i < 32 && tid + i < NT; ++i) { P3 nx, r; load_pose(lane_tot + 7 * (tid + i), nx); se3_mul(acc, nx, r); acc = r; }
        store_pose(stage + 7 * w, acc);
    }
    __syncthreads();
    P3 acc; acc.t[0] = acc.t[1] = acc.t[2] = 0; acc.q[0] = 1; acc.q[1] = acc.q[2] = acc.q[3] = 0;
    for (int i = 0; i < w; ++i) { P3 nx, r; load_pose(stage + 7 * i, nx); se3_mul(acc, nx, r); acc = r; }
    for (int i = 0; i < lane; ++i) { P3 nx, r; load_pose(lane_tot + 7 * (w * 32 + i), nx); se3_mul(acc, nx, r); acc = r; }
    excl = acc;
    __syncthreads();
#else
    excl.t[0] = excl.t[1] = excl.t[2] = 0; excl.q[0] = 1; excl.q[1] = excl.q[2] = excl.q[3] = 0;
    if (NT > 1) {
        double* tot = M.st;
        const int t = hd_tid();
        store_pose(tot + 7 * t, mine);
        bsync<NT>();
        for (int i = 0; i < t; ++i) { P3 nx, r; load_pose(tot + 7 * i, nx); se3_mul(excl, nx, r); excl = r; }
        bsync<NT>();
    }
#endif
}

// One SE(3) check. odom_all: AoS records of ODOM_REC3 doubles per odometry edge (global index).
template <int NT> IPC_HD void run_check3(const ChainMem3& M_in, const double* odom_all, const LoopRec3* Lc_in, const LoopRec3* Lm_in, const CheckParams& prm,
                                         bool want_info, CheckResult& res) {
    const int tid = hd_tid();
    ChainMem3 M = M_in;                                     // st / stw swap when a trial is accepted (global-memory state); S is set below
    bsync<NT>();
    if (tid == 0) {
        UniBlock3* U = M.U();
        U->Lc = *Lc_in; U->Lm = Lm_in ? *Lm_in : *Lc_in;
        CheckGeom3 g;
        const int cf = U->Lc.from, ct = U->Lc.to;
        const int ca = cf < ct ? cf : ct, cb = cf < ct ? ct : cf;
        int lo = ca, hi = cb; g.K = 1;
        int ma = 0, mb = 0;
        if (Lm_in) {
            const int mf = U->Lm.from, mt = U->Lm.to;
            ma = mf < mt ? mf : mt; mb = mf < mt ? mt : mf;
            if ((mb < cb ? mb : cb) - (ma > ca ? ma : ca) > 0) { g.K = 2; lo = ca < ma ? ca : ma; hi = cb > mb ? cb : mb; }   // src/consensus.cpp:157-159
        }
        const int L = hi - lo;
        g.lo = lo; g.L = L;
        const int ca_l = ca - lo, cb_l = cb - lo;
        g.rs = 0; g.re = L; g.first_is_c = 1; g.last_is_c = 1;
        int ma_l = 0, mb_l = L;
        if (g.K == 2) {
            ma_l = ma - lo; mb_l = mb - lo;
            g.rs = ca_l > ma_l ? ca_l : ma_l; g.re = cb_l < mb_l ? cb_l : mb_l;
            g.first_is_c = (ca_l == 0); g.last_is_c = (cb_l == L);
        }
        g.c_a_is_rs = ca_l != 0; g.c_b_is_L = cb_l == L; g.m_a_is_rs = ma_l != 0; g.m_b_is_L = mb_l == L;
        U->g = g;
    }
    bsync<NT>();
    const int K = M.U()->g.K, L = M.U()->g.L;
    const int spec_v[NSPEC] = {0, M.U()->g.rs, M.U()->g.re, L};
    const double th = (K == 2) ? prm.slow_th : prm.fast_th;
    int max_iter = (K == 2) ? prm.slow_iter : prm.fast_iter;
    if (L + K > 100) max_iter *= 5;                        // src/consensus_utils.cpp:12-13
    const double* odom = odom_all + (size_t)ODOM_REC3 * M.U()->g.lo;

    ThreadState3 ts;
    int S = (L + NT - 1) / NT; if (S < 1) S = 1; S |= 1;
    ts.k0 = tid * S < L ? tid * S : L; ts.k1 = ts.k0 + S < L ? ts.k0 + S : L;
    const int k0 = ts.k0, k1 = ts.k1;
    M.S = S; M.nt = NT; M.cs = M.gst ? NT : 1;

    // ---- dead-reckoning (propagateGuess): compose the measurements. Records hold Z^-1: Z = (Z^-1)^-1.
    {
        P3 id; id.t[0] = id.t[1] = id.t[2] = 0; id.q[0] = 1; id.q[1] = id.q[2] = id.q[3] = 0;
        P3 mine = id;
        for (int k = k0; k < k1; ++k) {
            P3 zi; load_pose(odom + (size_t)ODOM_REC3 * k, zi);
            P3 zz, r; se3_rel(zi, id, zz); se3_mul(mine, zz, r); mine = r;
        }
        P3 excl; se3_excl_scan<NT>(M, mine, excl);
        if (tid == 0) { store_pose(M.P(0), id, M.cs); if (M.gst) store_pose(M.PW(0), id, M.cs); }
        ts.pa = excl;
        P3 cur = excl;
        for (int k = k0; k < k1; ++k) {
            P3 zi; load_pose(odom + (size_t)ODOM_REC3 * k, zi);
            P3 zz, r; se3_rel(zi, id, zz); se3_mul(cur, zz, r); q_normalize(r.q); cur = r;
            store_pose(M.Pit(k + 1, k - k0, tid), cur, M.cs);
        }
#pragma unroll
        for (int m = 0; m < NP3; ++m) ts.base[m] = 0;
    }

    int buf = 0, n_sweeps = 0;
    SweepOut3 so;
    double cur_chi = 0, cur_max = 0, cand_chi = 0;
    double delta = 1e4;
    int iterations = 0, evals = 0, it = 0, tries = 0;
    double prev_hnorm = -1;
    bool have_norm = false, have_sd = false, need_rollback = false;
    bool last_bound = false;     // the previous decision was trust-region bound
    double hgnNorm = 0, bb = 0, bh = 0, hh = 0, bHb = 0, alpha = 0, hsdNorm = 0, linearGain = 0;
    double gain_loops = 0, gn_gain_model = 0;
    bool acc_gain = false;
    int purpose = P_INIT, mode = STEP_NONE;
    double c1 = 0, c2 = 0;
    for (;;) {
        if (need_rollback) { rollback3<NT>(M, ts); need_rollback = false; }
        sweep3<NT>(M, odom, mode, c1, c2, acc_gain && mode == STEP_GN, ts, so, buf, spec_v); ++n_sweeps;
        bool start_iter = false, after_reject = false, decide = false;
        if (purpose == P_INIT) {
            double n_c, n_m;
            eval_and_solve3<NT>(M, buf, so.chi, 0, 1, true, n_c, n_m);
            cur_chi = so.chi + n_c + n_m; cur_max = fmax(so.mx, fmax(n_c, n_m)); cand_chi = n_c;
            gain_loops = M.U()->sol.gain_loops;
            start_iter = true;
        } else if (purpose == P_RELIN_SPECFAIL) {
            decide = true;
        } else if (purpose == P_RELIN_REJECT) {
            after_reject = true;
        } else {
            bool specfail = false;
            if (purpose == P_TRIAL_SPEC) { hgnNorm = sqrt(so.hh); have_norm = true; specfail = !(hgnNorm < delta); }
            if (specfail) { need_rollback = true; mode = STEP_NONE; purpose = P_RELIN_SPECFAIL; continue; }
            const bool trial_gn = purpose != P_TRIAL_BLEND;
            const double hdlNorm = sqrt(so.hh);
            if (trial_gn) linearGain = acc_gain ? so.gain + gain_loops : gn_gain_model;
            ++evals;
            double n_c, n_m;
            eval_and_solve3<NT>(M, buf, so.chi, cur_chi, linearGain, false, n_c, n_m);
            const double newChi = so.chi + n_c + n_m;
            const double rawGain = linearGain;
            double lg = linearGain;
            if (fabs(lg) < 1e-12) lg = 1e-12;
            const double rho = (cur_chi - newChi) / lg;
            if (rho > 0.75) delta = fmax(delta, 3 * hdlNorm);
            else if (rho < 0.25) delta *= 0.5;
            if (rho > 0) {
                if (M.gst) { double* t_ = M.st; M.st = M.stw; M.stw = t_; }     // the trial state becomes the state
                cur_chi = newChi; cur_max = fmax(so.mx, fmax(n_c, n_m)); cand_chi = n_c;
                gain_loops = M.U()->sol.gain_loops;
                prev_hnorm = hdlNorm;
                ++iterations; ++it;
                start_iter = true;
            } else {
                need_rollback = true;
                prev_hnorm = -1;
                if (trial_gn) while (tries < prm.max_tries && hgnNorm < delta) { ++tries; ++evals; delta *= 0.5; }
                if (prm.noise_eps > 0 && rawGain <= prm.noise_eps * cur_chi + 1e-300) tries = prm.max_tries;
                if (trial_gn && tries < prm.max_tries) { mode = STEP_NONE; purpose = P_RELIN_REJECT; continue; }
                after_reject = true;
            }
        }
        if (after_reject) {
            if (tries < prm.max_tries) decide = true;
            else { ++iterations; break; }
            ++tries;
        }
        if (start_iter) {
            if (it >= max_iter) break;
            if (prm.early_accept && !want_info && cur_chi <= th) break;
            have_norm = false; have_sd = false; tries = 1;
            decide = true;
        }
        if (decide) {
            if (!have_norm) {
                gn_gain_model = cur_chi - M.U()->sol.model;
                acc_gain = true;   // always accumulate (the SE(3) edge loop is dominated by the 6x6 algebra)
                if (prm.speculate && prev_hnorm >= 0 && 4 * prev_hnorm < delta) {
                    mode = STEP_GN; purpose = P_TRIAL_SPEC; c1 = 0; c2 = 1;
                    continue;
                }
                if (need_rollback) { rollback3<NT>(M, ts); need_rollback = false; }
                if (prm.sd_fuse && last_bound) {    // expected trust-region bound: the gradient passes also deliver |h_gn|^2
                    sd_sweeps3<NT>(M, odom, ts, bb, bh, hh, bHb); n_sweeps += 2;
                    hgnNorm = sqrt(hh);
                    alpha = bb / bHb; hsdNorm = alpha * sqrt(bb); have_sd = true;
                } else {
                    hgnNorm = sqrt(gn_norm_sq3<NT>(M, odom, ts)); ++n_sweeps;
                }
                have_norm = true;
            }
            if (hgnNorm < delta) {
                mode = STEP_GN; purpose = P_TRIAL_GN; c1 = 0; c2 = 1; last_bound = false;
                continue;
            }
            last_bound = true;
            if (!have_sd) {
                if (need_rollback) { rollback3<NT>(M, ts); need_rollback = false; }
                sd_sweeps3<NT>(M, odom, ts, bb, bh, hh, bHb); n_sweeps += 2;
                alpha = bb / bHb;
                hsdNorm = alpha * sqrt(bb);
                have_sd = true;
            }
            if (hsdNorm > delta) { c1 = delta / hsdNorm * alpha; c2 = 0; }
            else {
                const double hsdSq = alpha * alpha * bb;
                const double c = alpha * bh - hsdSq;
                const double bma = hh - 2 * alpha * bh + hsdSq;
                double beta;
                if (c <= 0) beta = (-c + sqrt(c * c + bma * (delta * delta - hsdSq))) / bma;
                else beta = (delta * delta - hsdSq) / (c + sqrt(c * c + bma * (delta * delta - hsdSq)));
                c1 = alpha * (1 - beta); c2 = beta;
            }
            linearGain = -(c1 * c1 * bHb + 2 * c1 * c2 * bb + c2 * c2 * bh) + 2 * (c1 * bb + c2 * bh);
            mode = STEP_BLEND; purpose = P_TRIAL_BLEND;
            continue;
        }
        break;
    }
    res.verdict = (cur_max > th) ? 0 : 1;
    res.max_chi2 = cur_max; res.cand_chi2 = cand_chi; res.sum_chi2 = cur_chi;
    res.iterations = iterations; res.evals = evals; res.window_len = L; res.n_loops = K; res.n_sweeps = n_sweeps;
}

}  // namespace se3
}  // namespace ipcb
