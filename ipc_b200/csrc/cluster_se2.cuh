// cluster_se2.cuh — SE(2) window solve with ANY number of loop edges (the slow path of the sequential stream): the per-phase
// device functions of the persistent stream solver (stream_solver.cuh).
//
// Replaces isAgreeingWithCurrentState (/root/reference/src/consensus_utils.cpp:6-22) as driven by
// IPC::agreementCheck (/root/reference/src/consensus.cpp:42-75) when the candidate's cluster holds K - 1 >= 0 accepted
// loops. Same twist-coordinate formulation as chain_se2.cuh (DESIGN.md "Chain solve"), generalised:
//   * per odometry edge k: terms M_k = Q_k V_k Q_k^T, m_k = -Q_k d_k; prefix sums PM, Pm over the window;
//   * per loop l on the edge interval [a_l, b_l): W_l = Q_l V_l Q_l^T, sigma_l, d_l;
//   * forces: (P(I_l ∩ I_l') + delta_ll' W_l) z_l' = Pm(b_l) - Pm(a_l) + sigma_l Q_l d_l  — dense SPD, 3K x 3K, factorised by the
//     hand-written blocked Cholesky of stream_solver.cuh (no library call anywhere on this path);
//   * step: f_k = sum of z_l over the loops covering edge k, xi_k = m_k - M_k f_k, Xi = prefix(xi), h_j = T_j Xi_j.
// Every function here is executed by ONE CTA of CL_NT threads (CTA 0 of the cooperative grid) unless it says "grid".
#pragma once
#include "chain_se2.cuh"

namespace ipcb {

constexpr int CL_NT = 512;      // threads per CTA of the stream solver (128 registers per thread)
constexpr int CL_NRES = 16;     // doubles in the result buffer

struct ClLoop {                 // one loop edge of the cluster, local indices
    int jf, jt;                 // local vertex indices of from / to
    int a, b;                   // edge interval [a, b)
    double meas[3], D[6], V[6];
};

struct ClBuffers {
    double* W;                  // window state AoS[5] x (L + 1): x y theta cos sin  (SE(3): AoS[7], t + quaternion)
    double* T;                  // per-edge terms, SoA [NPRE][Lcap]
    double* P;                  // inclusive prefix per vertex, SoA [NPRE][Lcap + 1] (P[.][0] = 0)
    double* chi_e;              // per-edge chi2 [Lcap]
    double* lt;                 // per-loop terms [K][12]: t(9), sigma, chi, pad   (SE(3): [K][32])
};

// Loop end points by window position (CSR over local vertices 0..L, built on the host once per check): entry = loop index l
// shifted left by one, low bit set when the position is the START a_l of the loop's interval (else its END b_l). Within a
// position the entries are in increasing l: per-vertex sums run in loop order, i.e. deterministic.
struct ClEvents { const int* ptr; const int* idx; };

template <int M> __device__ __forceinline__ void cl_block_sum(double* v, double* red) {
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[m] += __shfl_xor_sync(0xffffffffu, v[m], o);
    const int w = threadIdx.x >> 5, NW = blockDim.x / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int m = 0; m < M; ++m) red[w * M + m] = v[m];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M; ++m) { double s = 0; for (int i = 0; i < NW; ++i) s += red[i * M + m]; v[m] = s; }
}
__device__ __forceinline__ double cl_block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, NW = blockDim.x / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = red[0];
    for (int i = 1; i < NW; ++i) s = fmax(s, red[i]);
    return s;
}
template <int M> __device__ __forceinline__ void cl_block_excl_scan(double* v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double inc[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double x = v[m];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        inc[m] = x;
    }
    __syncthreads();
    if (lane == 31)
#pragma unroll
        for (int m = 0; m < M; ++m) red[w * M + m] = inc[m];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M; ++m) { double base = 0; for (int i = 0; i < w; ++i) base += red[i * M + m]; v[m] = base + inc[m] - v[m]; }
}

// block-strided copy (one CTA)
__device__ __forceinline__ void cl_copy_cta(const double* src, double* dst, long long n) {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// linearise every odometry edge of the window: terms, prefix sums, chi2. res[0] = sum chi2, res[1] = max chi2.
__device__ __noinline__ void cl_linearize(const double* __restrict__ odom9, int lo, int L, int Lcap, ClBuffers B, double* res, double* red) {
    const int S = (L + CL_NT - 1) / CL_NT;
    const int k0 = min((int)threadIdx.x * S, L), k1 = min(k0 + S, L);
    double run[NPRE];
#pragma unroll
    for (int m = 0; m < NPRE; ++m) run[m] = 0;
    double chi = 0, mx = 0;
    for (int k = k0; k < k1; ++k) {
        const double* pa = B.W + 5 * k; const double* pb = pa + 5;
        const double* r = odom9 + 9 * (size_t)(lo + k);
        double V[6]; inv_sym3(r + 3, V);
        Lin2 e; lin2cs(pa[3], pa[4], P2{pa[0], pa[1], pa[2]}, P2{pb[0], pb[1], pb[2]}, r[0], r[1], r[2], r + 3, e);
        double t[NPRE]; edge_prefix_terms(e, V, pb[0], pb[1], t);
#pragma unroll
        for (int m = 0; m < NPRE; ++m) { B.T[(size_t)m * Lcap + k] = t[m]; run[m] += t[m]; }
        B.chi_e[k] = e.chi; chi += e.chi; mx = fmax(mx, e.chi);
    }
    cl_block_excl_scan<NPRE>(run, red);          // run = prefix at vertex k0
    if (threadIdx.x == 0)
#pragma unroll
        for (int m = 0; m < NPRE; ++m) B.P[(size_t)m * (Lcap + 1)] = 0;
    for (int k = k0; k < k1; ++k)
#pragma unroll
        for (int m = 0; m < NPRE; ++m) { run[m] += B.T[(size_t)m * Lcap + k]; B.P[(size_t)m * (Lcap + 1) + k + 1] = run[m]; }
    double s[1] = {chi};
    cl_block_sum<1>(s, red);
    mx = cl_block_max(mx, red);
    if (threadIdx.x == 0) { res[0] = s[0]; res[1] = mx; }
    __syncthreads();
}

// loop edges at the window state: terms, sigma, chi2. res[2] = sum of loop chi2, res[3] = max, res[4] = chi2 of the LAST loop (the candidate)
__device__ __noinline__ void cl_loops(const ClLoop* __restrict__ loops, int K, ClBuffers B, double* res, double* red) {
    double chi = 0, mx = 0;
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const ClLoop& Lp = loops[l];
        const double* pf = B.W + 5 * Lp.jf; const double* pt = B.W + 5 * Lp.jt;
        Lin2 e; lin2cs(pf[3], pf[4], P2{pf[0], pf[1], pf[2]}, P2{pt[0], pt[1], pt[2]}, Lp.meas[0], Lp.meas[1], Lp.meas[2], Lp.D, e);
        double t[NPRE]; edge_prefix_terms(e, Lp.V, pt[0], pt[1], t);
        double* o = B.lt + 12 * (size_t)l;
#pragma unroll
        for (int m = 0; m < NPRE; ++m) o[m] = t[m];
        o[9] = Lp.jt > Lp.jf ? 1.0 : -1.0; o[10] = e.chi;
        chi += e.chi; mx = fmax(mx, e.chi);
        if (l == K - 1) res[4] = e.chi;
    }
    double s[1] = {chi};
    cl_block_sum<1>(s, red);
    mx = cl_block_max(mx, red);
    if (threadIdx.x == 0) { res[2] = s[0]; res[3] = mx; }
    __syncthreads();
}

// GRID: lower block triangle of S (3K x 3K, column-major with leading dimension ld) and the right-hand side, stored as the extra
// matrix row `rhs_row` (the factorisation then forward-substitutes it for free, stream_solver.cuh).
__device__ __forceinline__ void cl_assemble_grid(const ClLoop* __restrict__ loops, int K, int Lcap, ClBuffers B, double* Smat, int ld, int rhs_row, int grank, int gsize) {
    const long long total = (long long)K * (K + 1) / 2;
    for (long long idx = (long long)grank * blockDim.x + threadIdx.x; idx < total; idx += (long long)gsize * blockDim.x) {
        // idx -> (l, m) with l >= m: row block l, column block m
        int l = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
        while ((long long)l * (l + 1) / 2 > idx) --l;
        while ((long long)(l + 1) * (l + 2) / 2 <= idx) ++l;
        const int m = (int)(idx - (long long)l * (l + 1) / 2);
        const int a = max(loops[l].a, loops[m].a), b = min(loops[l].b, loops[m].b);
        double blk[6] = {0, 0, 0, 0, 0, 0};
        if (b > a) {
#pragma unroll
            for (int q = 0; q < 6; ++q) blk[q] = B.P[(size_t)q * (Lcap + 1) + b] - B.P[(size_t)q * (Lcap + 1) + a];
        }
        if (l == m) {
            const double* t = B.lt + 12 * (size_t)l;
#pragma unroll
            for (int q = 0; q < 6; ++q) blk[q] += t[q];
            const int la = loops[l].a, lb = loops[l].b;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                Smat[(size_t)(3 * l + q) * ld + rhs_row] = B.P[(size_t)(6 + q) * (Lcap + 1) + lb] - B.P[(size_t)(6 + q) * (Lcap + 1) + la] - t[9] * t[6 + q];
        }
        const double full[9] = {blk[0], blk[1], blk[2], blk[1], blk[3], blk[4], blk[2], blk[4], blk[5]};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Smat[(size_t)(3 * m + c) * ld + (3 * l + r)] = full[r * 3 + c];
    }
}

// force on edge k0 = sum of z_l over the loops covering it (loop order), then advanced edge by edge with the end-point events
__device__ __forceinline__ void cl_force_at(const ClLoop* __restrict__ loops, int K, const double* z, int k, double* f) {
    f[0] = f[1] = f[2] = 0;
    for (int l = 0; l < K; ++l)
        if (loops[l].a <= k && k < loops[l].b) { f[0] += z[3 * l]; f[1] += z[3 * l + 1]; f[2] += z[3 * l + 2]; }
}
__device__ __forceinline__ void cl_force_step(const ClEvents& E, const double* z, int k, double* f) {   // f(edge k-1) -> f(edge k)
    for (int q = E.ptr[k]; q < E.ptr[k + 1]; ++q) {
        const int l = E.idx[q] >> 1; const double sg = (E.idx[q] & 1) ? 1.0 : -1.0;
        f[0] += sg * z[3 * l]; f[1] += sg * z[3 * l + 1]; f[2] += sg * z[3 * l + 2];
    }
}

// GN step from the forces z: f_k = sum_{l covers k} z_l, xi_k = m_k - M_k f_k, Xi = prefix(xi), h_j = T_j Xi_j.
// H: AoS[3] x (L + 1). res[5] = |h|^2, res[6] = predicted gain h^T H h = sum over edges |J h|^2_Omega, accumulated edge by
// edge from non-negative terms (chi2 - model cancels catastrophically for gross outliers): an odometry edge contributes
// xi_k^T M_k^-1 xi_k, a loop eta^T W_l^-1 eta with eta = sigma W_l z_l - Q_l d_l.
__device__ __noinline__ void cl_gn_step(const ClLoop* __restrict__ loops, int K, ClEvents E, int L, int Lcap, ClBuffers B, const double* z,
                                        double* H, double* res, double* red) {
    const int S = (L + CL_NT - 1) / CL_NT;
    const int k0 = min((int)threadIdx.x * S, L), k1 = min(k0 + S, L);
    double tot[3] = {0, 0, 0}, model = 0;
    double f[3] = {0, 0, 0};
    if (k0 < k1) cl_force_at(loops, K, z, k0, f);
    for (int k = k0; k < k1; ++k) {
        if (k > k0) cl_force_step(E, z, k, f);
        double t[NPRE];
#pragma unroll
        for (int m = 0; m < NPRE; ++m) t[m] = B.T[(size_t)m * Lcap + k];
        double Mf[3]; sym3_mul(t, f, Mf);
        const double xi[3] = {t[6] - Mf[0], t[7] - Mf[1], t[8] - Mf[2]};
        double Mi[6]; inv_sym3(t, Mi);
        model += quad3(Mi, xi[0], xi[1], xi[2]);
        tot[0] += xi[0]; tot[1] += xi[1]; tot[2] += xi[2];
        H[3 * (k + 1)] = tot[0]; H[3 * (k + 1) + 1] = tot[1]; H[3 * (k + 1) + 2] = tot[2];     // local inclusive prefix
    }
    cl_block_excl_scan<3>(tot, red);
    double hh = 0;
    for (int k = k0; k < k1; ++k) {
        const int j = k + 1;
        const double gx = H[3 * j] + tot[0], gy = H[3 * j + 1] + tot[1], gt = H[3 * j + 2] + tot[2];
        const double x = B.W[5 * j], y = B.W[5 * j + 1];
        const double hx = gx - y * gt, hy = gy + x * gt;
        H[3 * j] = hx; H[3 * j + 1] = hy; H[3 * j + 2] = gt;
        hh += hx * hx + hy * hy + gt * gt;
    }
    if (threadIdx.x == 0) { H[0] = 0; H[1] = 0; H[2] = 0; }
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const double* t = B.lt + 12 * (size_t)l;
        double Wz[3]; sym3_mul(t, z + 3 * l, Wz);
        const double eta[3] = {t[9] * Wz[0] + t[6], t[9] * Wz[1] + t[7], t[9] * Wz[2] + t[8]};
        double Wi[6]; inv_sym3(t, Wi);
        model += quad3(Wi, eta[0], eta[1], eta[2]);
    }
    double s[2] = {hh, model};
    cl_block_sum<2>(s, red);
    if (threadIdx.x == 0) { res[5] = s[0]; res[6] = s[1]; }
    __syncthreads();
}

// gradient b_j in g2o vertex coordinates, G: AoS[3] x (L + 1): odometry part, then the loop edges incident to j in loop order.
// lg: per-loop staging [K][6] (gi, gj of every loop at the current state).
__device__ __noinline__ void cl_gradient(const double* __restrict__ odom9, const ClLoop* __restrict__ loops, int K, ClEvents E, int lo, int L, ClBuffers B,
                                         double* G, double* lg) {
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const ClLoop& Lp = loops[l];
        const double* pf = B.W + 5 * Lp.jf; const double* pt = B.W + 5 * Lp.jt;
        Lin2 e; lin2cs(pf[3], pf[4], P2{pf[0], pf[1], pf[2]}, P2{pt[0], pt[1], pt[2]}, Lp.meas[0], Lp.meas[1], Lp.meas[2], Lp.D, e);
        grad2(e, lg + 6 * (size_t)l, lg + 6 * (size_t)l + 3);
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
        double b[3] = {0, 0, 0};
        if (j > 0) {
            const double* pa = B.W + 5 * (j - 1); const double* pb = pa + 5;
            const double* r = odom9 + 9 * (size_t)(lo + j - 1);
            Lin2 e; lin2cs(pa[3], pa[4], P2{pa[0], pa[1], pa[2]}, P2{pb[0], pb[1], pb[2]}, r[0], r[1], r[2], r + 3, e);
            double gi[3], gj[3]; grad2(e, gi, gj);
            b[0] -= gj[0]; b[1] -= gj[1]; b[2] -= gj[2];
            if (j < L) {
                const double* pc = pb + 5;
                const double* r2 = r + 9;
                Lin2 e2; lin2cs(pb[3], pb[4], P2{pb[0], pb[1], pb[2]}, P2{pc[0], pc[1], pc[2]}, r2[0], r2[1], r2[2], r2 + 3, e2);
                grad2(e2, gi, gj);
                b[0] -= gi[0]; b[1] -= gi[1]; b[2] -= gi[2];
            }
            for (int q = E.ptr[j]; q < E.ptr[j + 1]; ++q) {
                const int l = E.idx[q] >> 1;
                const double* g = lg + 6 * (size_t)l + (loops[l].jf == j ? 0 : 3);
                b[0] -= g[0]; b[1] -= g[1]; b[2] -= g[2];
            }
        }
        G[3 * j] = b[0]; G[3 * j + 1] = b[1]; G[3 * j + 2] = b[2];     // vertex 0 of the window is fixed: zero
    }
    __syncthreads();
}
// res[7] = |b|^2, res[8] = b . h_gn, res[9] = b^T H b
__device__ __noinline__ void cl_sd_scalars(const double* __restrict__ odom9, const ClLoop* __restrict__ loops, int K, int lo, int L, ClBuffers B,
                                           const double* G, const double* H, double* res, double* red) {
    double v[3] = {0, 0, 0};
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
        const double* b = G + 3 * j; const double* h = H + 3 * j;
        v[0] += b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
        v[1] += b[0] * h[0] + b[1] * h[1] + b[2] * h[2];
        if (j < L) {
            const double* pa = B.W + 5 * j; const double* pb = pa + 5;
            const double* r = odom9 + 9 * (size_t)(lo + j);
            Lin2 e; lin2cs(pa[3], pa[4], P2{pa[0], pa[1], pa[2]}, P2{pb[0], pb[1], pb[2]}, r[0], r[1], r[2], r + 3, e);
            double q0, q1, q2; dlin2(e, b, b + 3, q0, q1, q2);
            v[2] += quad3(r + 3, q0, q1, q2);
        }
    }
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const ClLoop& Lp = loops[l];
        const double* pf = B.W + 5 * Lp.jf; const double* pt = B.W + 5 * Lp.jt;
        Lin2 e; lin2cs(pf[3], pf[4], P2{pf[0], pf[1], pf[2]}, P2{pt[0], pt[1], pt[2]}, Lp.meas[0], Lp.meas[1], Lp.meas[2], Lp.D, e);
        double q0, q1, q2; dlin2(e, G + 3 * Lp.jf, G + 3 * Lp.jt, q0, q1, q2);
        v[2] += quad3(Lp.D, q0, q1, q2);
    }
    cl_block_sum<3>(v, red);
    if (threadIdx.x == 0) { res[7] = v[0]; res[8] = v[1]; res[9] = v[2]; }
    __syncthreads();
}
// trial state W1 = W0 (+) (c1 b + c2 h_gn); res[10] = |h|^2
__device__ __noinline__ void cl_apply(int L, const double* W0, const double* G, const double* H, double c1,
                                      double c2, double* W1, double* res, double* red) {
    double hh[1] = {0};
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
        double h[3] = {c2 * H[3 * j], c2 * H[3 * j + 1], c2 * H[3 * j + 2]};
        if (c1 != 0.0) { h[0] += c1 * G[3 * j]; h[1] += c1 * G[3 * j + 1]; h[2] += c1 * G[3 * j + 2]; }
        if (j == 0) { h[0] = h[1] = h[2] = 0; }
        const double* p = W0 + 5 * j; double* q = W1 + 5 * j;
        const double t = wrap_pi_hd(p[2] + h[2]);
        double s, c; ipc_sincos(t, &s, &c);
        q[0] = p[0] + h[0]; q[1] = p[1] + h[1]; q[2] = t; q[3] = c; q[4] = s;
        hh[0] += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
    }
    cl_block_sum<1>(hh, red);
    if (threadIdx.x == 0) res[10] = hh[0];
    __syncthreads();
}

// pose[j] for j = start+1 .. n-1 re-dead-reckoned from pose[start] (propagateCurrentGuess / propagateGuess,
// src/consensus_utils.cpp:60-71, 98-116): two block scans (headings, then rotated translations). One CTA of CL_NT threads.
__device__ __noinline__ void cl_dead_reckon_cta(const double* __restrict__ odom9, int start, int n, double* pose, double* red) {
    const int L = n - 1 - start;
    if (L <= 0) return;
    const int S = (L + CL_NT - 1) / CL_NT;
    const int k0 = min((int)threadIdx.x * S, L), k1 = min(k0 + S, L);
    const double th_s = pose[5 * (size_t)start + 2], x_s = pose[5 * (size_t)start], y_s = pose[5 * (size_t)start + 1];
    double v[1] = {0};
    for (int k = k0; k < k1; ++k) v[0] += odom9[9 * (size_t)(start + k) + 2];
    cl_block_excl_scan<1>(v, red);
    double acc = th_s + v[0];
    double thk = wrap_pi_hd(acc), s, c;
    ipc_sincos(thk, &s, &c);
    const double c_first = c, s_first = s;
    double p[2] = {0, 0};
    for (int k = k0; k < k1; ++k) {
        const double* r = odom9 + 9 * (size_t)(start + k);
        p[0] += c * r[0] - s * r[1]; p[1] += s * r[0] + c * r[1];
        acc += r[2]; thk = wrap_pi_hd(acc); ipc_sincos(thk, &s, &c);
        double* q = pose + 5 * (size_t)(start + k + 1);
        q[2] = thk; q[3] = c; q[4] = s;
    }
    cl_block_excl_scan<2>(p, red);
    double ax = x_s + p[0], ay = y_s + p[1];
    c = c_first; s = s_first;
    for (int k = k0; k < k1; ++k) {
        const double* r = odom9 + 9 * (size_t)(start + k);
        ax += c * r[0] - s * r[1]; ay += s * r[0] + c * r[1];
        double* q = pose + 5 * (size_t)(start + k + 1);
        q[0] = ax; q[1] = ay; c = q[3]; s = q[4];
    }
    __syncthreads();
}
__global__ void __launch_bounds__(CL_NT) cl_dead_reckon(const double* __restrict__ odom9, int start, int n, double* pose) {
    __shared__ double red[32 * 2];
    cl_dead_reckon_cta(odom9, start, n, pose, red);
}
__global__ void cl_set_origin(double* pose) { if (threadIdx.x == 0) { pose[0] = 0; pose[1] = 0; pose[2] = 0; pose[3] = 1; pose[4] = 0; } }
__global__ void cl_export_poses(const double* pose, int n, double* out) {
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += blockDim.x * gridDim.x) { out[3 * i] = pose[5 * i]; out[3 * i + 1] = pose[5 * i + 1]; out[3 * i + 2] = pose[5 * i + 2]; }
}

}  // namespace ipcb
