// cluster_se3.cuh — SE(3) window solve with any number of loop edges: the 6-dimensional copy of cluster_se2.cuh
// (sequential stream of IPC<EdgeSE3, VertexSE3>::agreementCheck, /root/reference/src/consensus.cpp:42-75,175) built on
// the device functions of chain_se3.cuh. Forces: (P(I_l ∩ I_l') + delta W_l) z_l' = Pm(b) - Pm(a) + sigma Q_l e_l, 6K x 6K
// dense SPD, factorised by the hand-written blocked Cholesky of stream_solver.cuh; the Dogleg runs on the device.
// Every function here is executed by ONE CTA of CL_NT threads (CTA 0 of the cooperative grid) unless it says "grid".
#pragma once
#include "chain_se3.cuh"
#include "cluster_se2.cuh"

namespace ipcb {

struct ClLoop3 {
    int jf, jt, a, b;
    double zinv[7], Om[se3::NS6], V[se3::NS6];
};
constexpr int CL3_LT = 32;      // per-loop terms: t(27), sigma, chi, pad

// res[0] = sum chi2, res[1] = max chi2 over the odometry edges
__device__ __noinline__ void cl3_linearize(const double* __restrict__ odom49, int lo, int L, int Lcap, ClBuffers B, double* res, double* red) {
    using namespace se3;
    const int S = (L + CL_NT - 1) / CL_NT;
    const int k0 = min((int)threadIdx.x * S, L), k1 = min(k0 + S, L);
    double run[NP3];
    for (int m = 0; m < NP3; ++m) run[m] = 0;
    double chi = 0, mx = 0;
    for (int k = k0; k < k1; ++k) {
        P3 pa, pb; load_pose(B.W + 7 * k, pa); load_pose(B.W + 7 * (k + 1), pb);
        const double* r = odom49 + (size_t)ODOM_REC3 * (lo + k);
        Lin3 e; lin3(r, pa, pb, r + 7, e);
        double t[NP3]; edge_terms3(e, r + 7 + NS6, t);
        for (int m = 0; m < NP3; ++m) { B.T[(size_t)m * Lcap + k] = t[m]; run[m] += t[m]; }
        B.chi_e[k] = e.chi; chi += e.chi; mx = fmax(mx, e.chi);
    }
    cl_block_excl_scan<NP3>(run, red);
    if (threadIdx.x == 0) for (int m = 0; m < NP3; ++m) B.P[(size_t)m * (Lcap + 1)] = 0;
    for (int k = k0; k < k1; ++k)
        for (int m = 0; m < NP3; ++m) { run[m] += B.T[(size_t)m * Lcap + k]; B.P[(size_t)m * (Lcap + 1) + k + 1] = run[m]; }
    double s[1] = {chi};
    cl_block_sum<1>(s, red);
    mx = cl_block_max(mx, red);
    if (threadIdx.x == 0) { res[0] = s[0]; res[1] = mx; }
    __syncthreads();
}

__device__ __noinline__ void cl3_loops(const ClLoop3* __restrict__ loops, int K, ClBuffers B, double* res, double* red) {
    using namespace se3;
    double chi = 0, mx = 0;
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const ClLoop3& Lp = loops[l];
        P3 pf, pt; load_pose(B.W + 7 * Lp.jf, pf); load_pose(B.W + 7 * Lp.jt, pt);
        Lin3 e; lin3(Lp.zinv, pf, pt, Lp.Om, e);
        double t[NP3]; edge_terms3(e, Lp.V, t);
        double* o = B.lt + CL3_LT * (size_t)l;
        for (int m = 0; m < NP3; ++m) o[m] = t[m];
        o[27] = Lp.jt > Lp.jf ? 1.0 : -1.0; o[28] = e.chi;
        chi += e.chi; mx = fmax(mx, e.chi);
        if (l == K - 1) res[4] = e.chi;
    }
    double s[1] = {chi};
    cl_block_sum<1>(s, red);
    mx = cl_block_max(mx, red);
    if (threadIdx.x == 0) { res[2] = s[0]; res[3] = mx; }
    __syncthreads();
}

// GRID: lower block triangle of S (6K x 6K, column-major, leading dimension ld) + the right-hand side as matrix row `rhs_row`
__device__ __forceinline__ void cl3_assemble_grid(const ClLoop3* __restrict__ loops, int K, int Lcap, ClBuffers B, double* Smat, int ld, int rhs_row, int grank, int gsize) {
    using namespace se3;
    const long long total = (long long)K * (K + 1) / 2;
    for (long long idx = (long long)grank * blockDim.x + threadIdx.x; idx < total; idx += (long long)gsize * blockDim.x) {
        int l = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
        while ((long long)l * (l + 1) / 2 > idx) --l;
        while ((long long)(l + 1) * (l + 2) / 2 <= idx) ++l;
        const int m = (int)(idx - (long long)l * (l + 1) / 2);
        const int a = max(loops[l].a, loops[m].a), b = min(loops[l].b, loops[m].b);
        double blk[NS6];
        for (int q = 0; q < NS6; ++q) blk[q] = (b > a) ? B.P[(size_t)q * (Lcap + 1) + b] - B.P[(size_t)q * (Lcap + 1) + a] : 0.0;
        if (l == m) {
            const double* t = B.lt + CL3_LT * (size_t)l;
            for (int q = 0; q < NS6; ++q) blk[q] += t[q];
            const int la = loops[l].a, lb = loops[l].b;
            for (int q = 0; q < 6; ++q)
                Smat[(size_t)(6 * l + q) * ld + rhs_row] = B.P[(size_t)(NS6 + q) * (Lcap + 1) + lb] - B.P[(size_t)(NS6 + q) * (Lcap + 1) + la] - t[27] * t[NS6 + q];
        }
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) Smat[(size_t)(6 * m + c) * ld + (6 * l + r)] = blk[sidx(r, c)];
    }
}

__device__ __forceinline__ void cl3_force_at(const ClLoop3* __restrict__ loops, int K, const double* z, int k, double* f) {
    for (int q = 0; q < 6; ++q) f[q] = 0;
    for (int l = 0; l < K; ++l)
        if (loops[l].a <= k && k < loops[l].b) for (int q = 0; q < 6; ++q) f[q] += z[6 * l + q];
}
__device__ __forceinline__ void cl3_force_step(const ClEvents& E, const double* z, int k, double* f) {
    for (int e = E.ptr[k]; e < E.ptr[k + 1]; ++e) {
        const int l = E.idx[e] >> 1; const double sg = (E.idx[e] & 1) ? 1.0 : -1.0;
        for (int q = 0; q < 6; ++q) f[q] += sg * z[6 * l + q];
    }
}

// H: AoS[6] x (L + 1), g2o vertex coordinates. res[5] = |h|^2, res[6] = predicted gain (sum of xi^T M^-1 xi and loop terms)
__device__ __noinline__ void cl3_gn_step(const ClLoop3* __restrict__ loops, int K, ClEvents E, int L, int Lcap, ClBuffers B, const double* z,
                                         double* H, double* res, double* red) {
    using namespace se3;
    const int S = (L + CL_NT - 1) / CL_NT;
    const int k0 = min((int)threadIdx.x * S, L), k1 = min(k0 + S, L);
    double tot[6] = {0, 0, 0, 0, 0, 0}, gain = 0;
    double f[6] = {0, 0, 0, 0, 0, 0};
    if (k0 < k1) cl3_force_at(loops, K, z, k0, f);
    for (int k = k0; k < k1; ++k) {
        if (k > k0) cl3_force_step(E, z, k, f);
        double t[NP3];
        for (int m = 0; m < NP3; ++m) t[m] = B.T[(size_t)m * Lcap + k];
        double Mf[6]; sym6_vec(t, f, Mf);
        double xi[6];
        for (int q = 0; q < 6; ++q) xi[q] = t[NS6 + q] - Mf[q];
        {   // xi^T M^-1 xi
            double A[36], y[6];
            for (int r = 0; r < 6; ++r) { for (int c = 0; c < 6; ++c) A[r * 6 + c] = t[sidx(r, c)]; y[r] = xi[r]; }
            chol_solve<6>(A, y);
            for (int q = 0; q < 6; ++q) gain += xi[q] * y[q];
        }
        for (int q = 0; q < 6; ++q) { tot[q] += xi[q]; H[6 * (k + 1) + q] = tot[q]; }
    }
    cl_block_excl_scan<6>(tot, red);
    double hh = 0;
    for (int k = k0; k < k1; ++k) {
        const int j = k + 1;
        double Xi[6];
        for (int q = 0; q < 6; ++q) Xi[q] = H[6 * j + q] + tot[q];
        P3 p; load_pose(B.W + 7 * j, p);
        double R[9]; q_to_R(p.q, R);
        double c[3]; cross3(p.t, Xi + 3, c);
        const double a[3] = {Xi[0] - c[0], Xi[1] - c[1], Xi[2] - c[2]};
        double u[6], r[3];
        m3t_vec(R, a, u); m3t_vec(R, Xi + 3, r);
        u[3] = 0.5 * r[0]; u[4] = 0.5 * r[1]; u[5] = 0.5 * r[2];
        for (int q = 0; q < 6; ++q) { H[6 * j + q] = u[q]; hh += u[q] * u[q]; }
    }
    if (threadIdx.x == 0) for (int q = 0; q < 6; ++q) H[q] = 0;
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const double* t = B.lt + CL3_LT * (size_t)l;
        double Wz[6]; sym6_vec(t, z + 6 * l, Wz);
        double eta[6], A[36], y[6];
        for (int q = 0; q < 6; ++q) { eta[q] = t[27] * Wz[q] + t[NS6 + q]; y[q] = eta[q]; }
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) A[r * 6 + c] = t[sidx(r, c)];
        chol_solve<6>(A, y);
        for (int q = 0; q < 6; ++q) gain += eta[q] * y[q];
    }
    double s[2] = {hh, gain};
    cl_block_sum<2>(s, red);
    if (threadIdx.x == 0) { res[5] = s[0]; res[6] = s[1]; }
    __syncthreads();
}

// gradient b_j, G: AoS[6] x (L + 1): odometry part, then the loop edges incident to j in loop order. lg: staging [K][12]
__device__ __noinline__ void cl3_gradient(const double* __restrict__ odom49, const ClLoop3* __restrict__ loops, int K, ClEvents E, int lo, int L, ClBuffers B,
                                          double* G, double* lg) {
    using namespace se3;
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const ClLoop3& Lp = loops[l];
        P3 pf, pt; load_pose(B.W + 7 * Lp.jf, pf); load_pose(B.W + 7 * Lp.jt, pt);
        Lin3 e; lin3(Lp.zinv, pf, pt, Lp.Om, e);
        double Ji[36], Jj[36]; jac3(Lp.zinv, e, Ji, Jj);
        m6t_vec(Ji, e.we, lg + 12 * (size_t)l); m6t_vec(Jj, e.we, lg + 12 * (size_t)l + 6);
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
        double b[6] = {0, 0, 0, 0, 0, 0};
        if (j > 0) {
            P3 pa, pb; load_pose(B.W + 7 * (j - 1), pa); load_pose(B.W + 7 * j, pb);
            const double* r = odom49 + (size_t)ODOM_REC3 * (lo + j - 1);
            Lin3 e; lin3(r, pa, pb, r + 7, e);
            double Ji[36], Jj[36], g[6]; jac3(r, e, Ji, Jj);
            m6t_vec(Jj, e.we, g);
            for (int q = 0; q < 6; ++q) b[q] -= g[q];
            if (j < L) {
                P3 pc; load_pose(B.W + 7 * (j + 1), pc);
                const double* r2 = r + ODOM_REC3;
                Lin3 e2; lin3(r2, pb, pc, r2 + 7, e2);
                jac3(r2, e2, Ji, Jj);
                m6t_vec(Ji, e2.we, g);
                for (int q = 0; q < 6; ++q) b[q] -= g[q];
            }
            for (int e2 = E.ptr[j]; e2 < E.ptr[j + 1]; ++e2) {
                const int l = E.idx[e2] >> 1;
                const double* gl = lg + 12 * (size_t)l + (loops[l].jf == j ? 0 : 6);
                for (int q = 0; q < 6; ++q) b[q] -= gl[q];
            }
        }
        for (int q = 0; q < 6; ++q) G[6 * j + q] = b[q];
    }
    __syncthreads();
}
// res[7] = |b|^2, res[8] = b . h_gn, res[9] = b^T H b
__device__ __noinline__ void cl3_sd_scalars(const double* __restrict__ odom49, const ClLoop3* __restrict__ loops, int K, int lo, int L, ClBuffers B,
                                            const double* G, const double* H, double* res, double* red) {
    using namespace se3;
    double v[3] = {0, 0, 0};
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
        const double* b = G + 6 * j; const double* h = H + 6 * j;
        for (int q = 0; q < 6; ++q) { v[0] += b[q] * b[q]; v[1] += b[q] * h[q]; }
        if (j < L) {
            P3 pa, pb; load_pose(B.W + 7 * j, pa); load_pose(B.W + 7 * (j + 1), pb);
            const double* r = odom49 + (size_t)ODOM_REC3 * (lo + j);
            Lin3 e; lin3(r, pa, pb, r + 7, e);
            double Ji[36], Jj[36], q1[6], q2[6]; jac3(r, e, Ji, Jj);
            m6_vec(Ji, b, q1); m6_vec(Jj, b + 6, q2);
            for (int q = 0; q < 6; ++q) q1[q] += q2[q];
            v[2] += sym6_quad(r + 7, q1);
        }
    }
    for (int l = threadIdx.x; l < K; l += blockDim.x) {
        const ClLoop3& Lp = loops[l];
        P3 pf, pt; load_pose(B.W + 7 * Lp.jf, pf); load_pose(B.W + 7 * Lp.jt, pt);
        Lin3 e; lin3(Lp.zinv, pf, pt, Lp.Om, e);
        double Ji[36], Jj[36], q1[6], q2[6]; jac3(Lp.zinv, e, Ji, Jj);
        m6_vec(Ji, G + 6 * Lp.jf, q1); m6_vec(Jj, G + 6 * Lp.jt, q2);
        for (int q = 0; q < 6; ++q) q1[q] += q2[q];
        v[2] += sym6_quad(Lp.Om, q1);
    }
    cl_block_sum<3>(v, red);
    if (threadIdx.x == 0) { res[7] = v[0]; res[8] = v[1]; res[9] = v[2]; }
    __syncthreads();
}
__device__ __noinline__ void cl3_apply(int L, const double* W0, const double* G, const double* H, double c1,
                                       double c2, double* W1, double* res, double* red) {
    using namespace se3;
    double hh[1] = {0};
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
        double u[6];
        for (int q = 0; q < 6; ++q) { u[q] = c2 * H[6 * j + q]; if (c1 != 0.0) u[q] += c1 * G[6 * j + q]; if (j == 0) u[q] = 0; }
        P3 p, o; load_pose(W0 + 7 * j, p);
        oplus3(p, u, o);
        store_pose(W1 + 7 * j, o);
        for (int q = 0; q < 6; ++q) hh[0] += u[q] * u[q];
    }
    cl_block_sum<1>(hh, red);
    if (threadIdx.x == 0) res[10] = hh[0];
    __syncthreads();
}
// pose[j], j = start+1 .. n-1, re-dead-reckoned from pose[start]; `stage` holds CL_NT x 7 doubles (global). One CTA of CL_NT threads.
__device__ __noinline__ void cl3_dead_reckon_cta(const double* __restrict__ odom49, int start, int n, double* pose, double* stage) {
    using namespace se3;
    const int L = n - 1 - start;
    if (L <= 0) return;
    const int S = (L + CL_NT - 1) / CL_NT;
    const int k0 = min((int)threadIdx.x * S, L), k1 = min(k0 + S, L);
    P3 id; id.t[0] = id.t[1] = id.t[2] = 0; id.q[0] = 1; id.q[1] = id.q[2] = id.q[3] = 0;
    P3 mine = id;
    for (int k = k0; k < k1; ++k) {
        P3 zi; load_pose(odom49 + (size_t)ODOM_REC3 * (start + k), zi);
        P3 zz, r; se3_rel(zi, id, zz); se3_mul(mine, zz, r); mine = r;
    }
    store_pose(stage + 7 * threadIdx.x, mine);
    __syncthreads();
    if (threadIdx.x == 0) {   // sequential exclusive scan of the thread totals, seeded with pose[start]
        P3 acc; load_pose(pose + 7 * (size_t)start, acc);
        for (int t = 0; t < CL_NT; ++t) {
            P3 tt, r; load_pose(stage + 7 * t, tt);
            store_pose(stage + 7 * t, acc);
            se3_mul(acc, tt, r); q_normalize(r.q); acc = r;
        }
    }
    __syncthreads();
    P3 cur; load_pose(stage + 7 * threadIdx.x, cur);
    for (int k = k0; k < k1; ++k) {
        P3 zi; load_pose(odom49 + (size_t)ODOM_REC3 * (start + k), zi);
        P3 zz, r; se3_rel(zi, id, zz); se3_mul(cur, zz, r); q_normalize(r.q); cur = r;
        store_pose(pose + 7 * (size_t)(start + k + 1), cur);
    }
    __syncthreads();
}
__global__ void __launch_bounds__(CL_NT) cl3_dead_reckon(const double* __restrict__ odom49, int start, int n, double* pose, double* stage) {
    cl3_dead_reckon_cta(odom49, start, n, pose, stage);
}
__global__ void cl3_set_origin(double* pose) { if (threadIdx.x == 0) { pose[0] = 0; pose[1] = 0; pose[2] = 0; pose[3] = 1; pose[4] = 0; pose[5] = 0; pose[6] = 0; } }
// (t, q = w x y z) -> g2o order x y z qx qy qz qw
__global__ void cl3_export_poses(const double* pose, int n, double* out) {
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += blockDim.x * gridDim.x) {
        const double* p = pose + 7 * (size_t)i; double* o = out + 7 * (size_t)i;
        const double sg = p[3] < 0 ? -1.0 : 1.0;      // same rotation, w >= 0 like g2o's writer
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = sg * p[4]; o[4] = sg * p[5]; o[5] = sg * p[6]; o[6] = sg * p[3];
    }
}

}  // namespace ipcb
