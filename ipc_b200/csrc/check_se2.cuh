// check_se2.cuh — batched SE(2) window check: one CTA per check (K = 1 or 2 loop edges).
//
// Replaces, for independent fast-path / pair sub-problems, the reference's
//   isAgreeingWithCurrentState  /root/reference/src/consensus_utils.cpp:6-22
// driven as in IPC::agreementCheck /root/reference/src/consensus.cpp:42-75 from the dead-reckoned
// state of IPC::IPC (:9-33), with g2o's Dogleg semantics.
//
// Formulation (DESIGN.md "Chain solve"): the Gauss-Newton system of the window is solved exactly, but
// not with a sparse factorisation. In the coordinates "global twist of odometry edge k"
//   xi_k = Q_k v_k,  Q_k = [R_k, -S t_{k+1}; 0 1],   u_j = T_j * sum_{k<j} xi_k,  T_j = [I, S t_j; 0 1]
// the odometry part of H is block diagonal and every loop edge is a low-rank term on an interval,
// so H^-1 b is: per-edge 3x3 work + interval sums (block reductions) + one 3x3 / 6x6 solve + a prefix
// sum. All Dogleg scalars (||h||, b^T h, h^T H h, rho, delta updates) are evaluated in g2o's own
// vertex coordinates u, so accept / reject decisions follow the reference.
#pragma once
#include "common.cuh"

namespace ipcb {

constexpr int SE2_NCOMP = 9;   // zx zy zt d00 d01 d02 d11 d12 d22

struct Pose2 { double x, y, t; };
struct Lin2 {            // linearisation of one relative-pose edge a -> b in its own frame
    double c, s;         // cos / sin of theta_a
    double rx, ry;       // R_a^T (t_b - t_a)
    double d0, d1, d2;   // residual in the relative frame (r - z)
    double w0, w1, w2;   // D d
    double chi;
};

__device__ __forceinline__ void lin2(const Pose2& a, const Pose2& b, double zx, double zy, double zt, const double* D, Lin2& e) {
    sincos(a.t, &e.s, &e.c);
    double dx = b.x - a.x, dy = b.y - a.y;
    e.rx = e.c * dx + e.s * dy;
    e.ry = -e.s * dx + e.c * dy;
    e.d0 = e.rx - zx; e.d1 = e.ry - zy; e.d2 = wrap_pi(b.t - a.t - zt);
    e.w0 = D[0] * e.d0 + D[1] * e.d1 + D[2] * e.d2;
    e.w1 = D[1] * e.d0 + D[3] * e.d1 + D[4] * e.d2;
    e.w2 = D[2] * e.d0 + D[4] * e.d1 + D[5] * e.d2;
    e.chi = e.d0 * e.w0 + e.d1 * e.w1 + e.d2 * e.w2;
}
__device__ __forceinline__ double quad3(const double* D, double a, double b, double c) {
    return a * (D[0] * a + D[1] * b + D[2] * c) + b * (D[1] * a + D[3] * b + D[4] * c) + c * (D[2] * a + D[4] * b + D[5] * c);
}
// linearised change of the edge residual under vertex increments ha (at a) and hb (at b), Jacobians at the old state
__device__ __forceinline__ void dlin2(const Lin2& e, const double* ha, const double* hb, double& q0, double& q1, double& q2) {
    double ux = hb[0] - ha[0], uy = hb[1] - ha[1];
    q0 = e.c * ux + e.s * uy + e.ry * ha[2];
    q1 = -e.s * ux + e.c * uy - e.rx * ha[2];
    q2 = hb[2] - ha[2];
}
// gradient pieces of chi2/2 w.r.t. the vertex increments: gi (at a), gj (at b)
__device__ __forceinline__ void grad2(const Lin2& e, double* gi, double* gj) {
    double rwx = e.c * e.w0 - e.s * e.w1, rwy = e.s * e.w0 + e.c * e.w1;   // R_a w_t
    gj[0] = rwx; gj[1] = rwy; gj[2] = e.w2;
    gi[0] = -rwx; gi[1] = -rwy; gi[2] = e.ry * e.w0 - e.rx * e.w1 - e.w2;
}
__device__ __forceinline__ void inv_sym3(const double* D, double* V) {
    double c00 = D[3] * D[5] - D[4] * D[4];
    double c01 = D[2] * D[4] - D[1] * D[5];
    double c02 = D[1] * D[4] - D[2] * D[3];
    double det = D[0] * c00 + D[1] * c01 + D[2] * c02;
    double id = 1.0 / det;
    V[0] = c00 * id; V[1] = c01 * id; V[2] = c02 * id;
    V[3] = (D[0] * D[5] - D[2] * D[2]) * id;
    V[4] = (D[1] * D[2] - D[0] * D[4]) * id;
    V[5] = (D[0] * D[3] - D[1] * D[1]) * id;
}
// dense n x n solve with partial pivoting (n <= 6), A row-major n x n, b overwritten by the solution
template <int N> __device__ __forceinline__ void solve_small(double* A, double* b) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
        int p = c; double best = fabs(A[c * N + c]);
#pragma unroll
        for (int r = c + 1; r < N; ++r) { double v = fabs(A[r * N + c]); if (v > best) { best = v; p = r; } }
        if (p != c) {
#pragma unroll
            for (int r = 0; r < N; ++r) if (r == p) {
#pragma unroll
                for (int k = 0; k < N; ++k) { double t = A[c * N + k]; A[c * N + k] = A[r * N + k]; A[r * N + k] = t; }
                double t = b[c]; b[c] = b[r]; b[r] = t;
            }
        }
        double inv = 1.0 / A[c * N + c];
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            double f = A[r * N + c] * inv;
#pragma unroll
            for (int k = c + 1; k < N; ++k) A[r * N + k] -= f * A[c * N + k];
            b[r] -= f * b[c];
        }
    }
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        double s = b[c];
#pragma unroll
        for (int k = c + 1; k < N; ++k) s -= A[c * N + k] * b[k];
        b[c] = s / A[c * N + c];
    }
}

struct LoopLin2 {
    int jf, jt;          // local vertex indices of from / to
    int hi_is_to;        // the larger-id endpoint is `to`
    int a, b;            // local edge interval [a, b)
    Lin2 e;
    double G[9];         // d(residual)/d(total twist over the interval)
    double g[3];         // G^T D d
    double Lam[6];       // G^T D G (sym)
};

// everything a thread needs about the loops of the check at the current state (computed redundantly per thread)
__device__ __forceinline__ void loop_lin2(const LoopRec2& L, int lo, const double* X, const double* Y, const double* TH, LoopLin2& o, bool need_G) {
    o.jf = L.from - lo; o.jt = L.to - lo;
    o.hi_is_to = L.to > L.from;
    o.a = min(o.jf, o.jt); o.b = max(o.jf, o.jt);
    Pose2 pf{X[o.jf], Y[o.jf], TH[o.jf]}, pt{X[o.jt], Y[o.jt], TH[o.jt]};
    lin2(pf, pt, L.meas[0], L.meas[1], L.meas[2], L.D, o.e);
    if (!need_G) return;
    const Lin2& e = o.e;
    double* G = o.G;
    if (o.hi_is_to) {   // G = [R_f^T, R_f^T S t_t; 0 1],  S t = (-y, x)
        G[0] = e.c; G[1] = e.s; G[2] = e.c * (-pt.y) + e.s * pt.x;
        G[3] = -e.s; G[4] = e.c; G[5] = -e.s * (-pt.y) + e.c * pt.x;
        G[6] = 0; G[7] = 0; G[8] = 1;
    } else {            // G = [-R_f^T, -R_f^T S t_f + (ry, -rx)^T; 0 -1]
        G[0] = -e.c; G[1] = -e.s; G[2] = -(e.c * (-pf.y) + e.s * pf.x) + e.ry;
        G[3] = e.s; G[4] = -e.c; G[5] = -(-e.s * (-pf.y) + e.c * pf.x) - e.rx;
        G[6] = 0; G[7] = 0; G[8] = -1;
    }
    // g = G^T w ; Lam = G^T D G
    o.g[0] = G[0] * e.w0 + G[3] * e.w1 + G[6] * e.w2;
    o.g[1] = G[1] * e.w0 + G[4] * e.w1 + G[7] * e.w2;
    o.g[2] = G[2] * e.w0 + G[5] * e.w1 + G[8] * e.w2;
    double DG[9];
    const double* D = L.D;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        DG[0 + c] = D[0] * G[c] + D[1] * G[3 + c] + D[2] * G[6 + c];
        DG[3 + c] = D[1] * G[c] + D[3] * G[3 + c] + D[4] * G[6 + c];
        DG[6 + c] = D[2] * G[c] + D[4] * G[3 + c] + D[5] * G[6 + c];
    }
    o.Lam[0] = G[0] * DG[0] + G[3] * DG[3] + G[6] * DG[6];
    o.Lam[1] = G[0] * DG[1] + G[3] * DG[4] + G[6] * DG[7];
    o.Lam[2] = G[0] * DG[2] + G[3] * DG[5] + G[6] * DG[8];
    o.Lam[3] = G[1] * DG[1] + G[4] * DG[4] + G[7] * DG[7];
    o.Lam[4] = G[1] * DG[2] + G[4] * DG[5] + G[7] * DG[8];
    o.Lam[5] = G[2] * DG[2] + G[5] * DG[5] + G[8] * DG[8];
}
__device__ __forceinline__ void sym3_mul(const double* S, const double* v, double* o) {
    o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}
// A(3x3 full) = P(sym) * L(sym)
__device__ __forceinline__ void sym3_sym3(const double* P, const double* L, double* A) {
    const double Pf[9] = {P[0], P[1], P[2], P[1], P[3], P[4], P[2], P[4], P[5]};
    const double Lf[9] = {L[0], L[1], L[2], L[1], L[3], L[4], L[2], L[4], L[5]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) A[r * 3 + c] = Pf[r * 3] * Lf[c] + Pf[r * 3 + 1] * Lf[3 + c] + Pf[r * 3 + 2] * Lf[6 + c];
}

// MODE 0: state + odometry records in shared memory (records staged by the TMA engine);
// MODE 1: state in shared memory, odometry records read straight from HBM/L2 (longer windows);
// MODE 2: state in a per-CTA global scratch as well (windows that exceed shared memory).
template <int NT, int MODE>
__global__ void __launch_bounds__(NT) check_chain_se2(BatchArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int capv = A.Lcap + 2;
    double* red = sm;                                   // (NT/32) * 32 doubles
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + (NT / 32) * 32);
    double* st = (MODE == 2) ? A.scratch + (size_t)blockIdx.x * 6 * capv : sm + (NT / 32) * 32 + 2;
    double* X = st;            double* Y = X + capv;   double* TH = Y + capv;
    double* HX = TH + capv;    double* HY = HX + capv; double* HT = HY + capv;
    double* CZ = HT + capv;    // MODE 0: 9 component arrays, capv each
    const int tid = threadIdx.x;
    const LoopRec2* loops = static_cast<const LoopRec2*>(A.loops);

    if (MODE == 0) {
        if (tid == 0) mbar_init(mbar, 1);
        __syncthreads();
    }
    uint32_t phase = 0;
    const int n_work = *A.n_work;

    for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const int chk = A.work[wi];
        const int cidx = A.cand[chk];
        int midx = A.member[chk];
        const LoopRec2 Lc = loops[cidx];
        LoopRec2 Lm = Lc;
        int ca = min(Lc.from, Lc.to), cb = max(Lc.from, Lc.to);
        int K = 1, lo = ca, hi = cb;
        if (midx >= 0) {
            Lm = loops[midx];
            int ma = min(Lm.from, Lm.to), mb = max(Lm.from, Lm.to);
            // src/consensus.cpp:157-159: positive-length overlap pulls the member into the cluster
            if (min(mb, cb) - max(ma, ca) > 0) { K = 2; lo = min(ca, ma); hi = max(cb, mb); }
        }
        const int L = hi - lo;
        const double th = (K == 2) ? A.slow_th : A.fast_th;
        int max_iter = (K == 2) ? A.slow_iter : A.fast_iter;
        if (L + K > 100) max_iter *= 5;                      // src/consensus_utils.cpp:12-13
        __syncthreads();                                     // previous check done with the arrays
        const double *ZX, *ZY, *ZT, *DD;
        int cstride;
        if (MODE == 0) {
            // ---- stage the window's odometry records (SoA) through the TMA engine ----------------
            const int sh = lo & 1;                           // 16-byte alignment of the bulk copies
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t nbytes = (uint32_t)(((L + sh + 1) & ~1) * 8);
                mbar_expect_tx(mbar, nbytes * SE2_NCOMP);
                for (int c = 0; c < SE2_NCOMP; ++c) bulk_g2s(CZ + c * capv, A.odom + (size_t)c * A.n_pad + (lo - sh), nbytes, mbar);
            }
            mbar_wait(mbar, phase);
            phase ^= 1;
            ZX = CZ + sh; ZY = CZ + capv + sh; ZT = CZ + 2 * capv + sh; DD = CZ + 3 * capv + sh;
            cstride = capv;
        } else {
            ZX = A.odom + lo; ZY = A.odom + A.n_pad + lo; ZT = A.odom + 2 * (size_t)A.n_pad + lo; DD = A.odom + 3 * (size_t)A.n_pad + lo;
            cstride = A.n_pad;
        }

        // segment of edges owned by this thread: odd length -> conflict-free strided smem access
        int S = (L + NT - 1) / NT; if (S < 1) S = 1; S |= 1;
        const int k0 = min(tid * S, L), k1 = min(k0 + S, L);

        // ---- dead-reckoning (propagateGuess, src/consensus_utils.cpp:98-116) as two block scans ----
        {
            double v[1] = {0};
            for (int k = k0; k < k1; ++k) v[0] += ZT[k];
            block_excl_scan<NT, 1>(v, red);
            double acc = v[0];
            if (tid == 0) { X[0] = 0; Y[0] = 0; TH[0] = 0; HX[0] = 0; HY[0] = 0; HT[0] = 0; }
            for (int k = k0; k < k1; ++k) { acc += ZT[k]; TH[k + 1] = wrap_pi(acc); }
            bsync<NT>();
            double p[2] = {0, 0};
            for (int k = k0; k < k1; ++k) { double s, c; sincos(TH[k], &s, &c); p[0] += c * ZX[k] - s * ZY[k]; p[1] += s * ZX[k] + c * ZY[k]; }
            block_excl_scan<NT, 2>(p, red);
            double ax = p[0], ay = p[1];
            for (int k = k0; k < k1; ++k) { double s, c; sincos(TH[k], &s, &c); ax += c * ZX[k] - s * ZY[k]; ay += s * ZX[k] + c * ZY[k]; X[k + 1] = ax; Y[k + 1] = ay; }
            bsync<NT>();
        }

        // interval structure (local edge indices): region0 = [0, rs) only `first`, region1 = [rs, re) all loops,
        // region2 = [re, L) only `last`
        const int ca_l = ca - lo, cb_l = cb - lo;
        int ma_l = 0, mb_l = 0, rs = 0, re = L, first_is_c = 1, last_is_c = 1;
        if (K == 2) {
            ma_l = min(Lm.from, Lm.to) - lo; mb_l = max(Lm.from, Lm.to) - lo;
            rs = max(ca_l, ma_l); re = min(cb_l, mb_l);
            first_is_c = (ca_l == 0); last_is_c = (cb_l == L);
        }

        // ---- Dogleg (OptimizationAlgorithmDogleg::solve + SparseOptimizer::optimize) --------------
        double delta = 1e4;
        int iterations = 0, evals = 0;
        double cur_chi = 0, cur_max = 0, cand_chi = 0;
        bool have_cur = false;          // cur_* valid for the current state
        bool ok = true;
        LoopLin2 lc, lm;

        for (int it = 0; it < max_iter && ok; ++it) {
            // ================= pass A: linearise, interval sums =================
            loop_lin2(Lc, lo, X, Y, TH, lc, true);
            if (K == 2) loop_lin2(Lm, lo, X, Y, TH, lm, true);
            double acc[3][9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 9; ++q) acc[r][q] = 0;
            double schi = 0, smax = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int rb = (r == 0) ? 0 : (r == 1 ? rs : re), rend = (r == 0) ? rs : (r == 1 ? re : L);
                const int ks = max(k0, rb), ke = min(k1, rend);
                if (ks >= ke) continue;
                Pose2 pa{X[ks], Y[ks], TH[ks]};
                for (int k = ks; k < ke; ++k) {
                    Pose2 pb{X[k + 1], Y[k + 1], TH[k + 1]};
                    double D[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) D[c] = DD[(size_t)c * cstride + k];
                    Lin2 e; lin2(pa, pb, ZX[k], ZY[k], ZT[k], D, e);
                    schi += e.chi; smax = fmax(smax, e.chi);
                    double V[6]; inv_sym3(D, V);
                    // Q = [c -s yb; s c -xb; 0 0 1]
                    const double q02 = pb.y, q12 = -pb.x;
                    double r0[3] = {e.c * V[0] - e.s * V[1] + q02 * V[2], e.c * V[1] - e.s * V[3] + q02 * V[4], e.c * V[2] - e.s * V[4] + q02 * V[5]};
                    double r1[3] = {e.s * V[0] + e.c * V[1] + q12 * V[2], e.s * V[1] + e.c * V[3] + q12 * V[4], e.s * V[2] + e.c * V[4] + q12 * V[5]};
                    acc[r][0] += r0[0] * e.c - r0[1] * e.s + r0[2] * q02;
                    acc[r][1] += r0[0] * e.s + r0[1] * e.c + r0[2] * q12;
                    acc[r][2] += r0[2];
                    acc[r][3] += r1[0] * e.s + r1[1] * e.c + r1[2] * q12;
                    acc[r][4] += r1[2];
                    acc[r][5] += V[5];
                    acc[r][6] += -(e.c * e.d0 - e.s * e.d1 + q02 * e.d2);
                    acc[r][7] += -(e.s * e.d0 + e.c * e.d1 + q12 * e.d2);
                    acc[r][8] += -e.d2;
                    pa = pb;
                }
            }
            {
                double v[29];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 9; ++q) v[r * 9 + q] = acc[r][q];
                v[27] = schi;
                v[28] = 0;
                block_sum<NT, 29>(v, red);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 9; ++q) acc[r][q] = v[r * 9 + q];
                schi = v[27];
            }
            smax = block_max<NT>(smax, red);
            cur_chi = schi + lc.e.chi + (K == 2 ? lm.e.chi : 0.0);
            cur_max = fmax(smax, fmax(lc.e.chi, K == 2 ? lm.e.chi : 0.0));
            cand_chi = lc.e.chi;
            have_cur = true;

            // ================= capacitance system: (I + P Lam) y = q - P g =================
            double zc[3], zm[3] = {0, 0, 0};   // z = g + Lam y  ("force" of each loop on its interval)
            if (K == 1) {
                double Amat[9], rhs[3], t[3];
                const double* P = acc[1];
                sym3_sym3(P, lc.Lam, Amat);
                Amat[0] += 1; Amat[4] += 1; Amat[8] += 1;
                sym3_mul(P, lc.g, t);
                rhs[0] = acc[1][6] - t[0]; rhs[1] = acc[1][7] - t[1]; rhs[2] = acc[1][8] - t[2];
                solve_small<3>(Amat, rhs);
                sym3_mul(lc.Lam, rhs, t);
                zc[0] = lc.g[0] + t[0]; zc[1] = lc.g[1] + t[1]; zc[2] = lc.g[2] + t[2];
            } else {
                // P_cc, P_mm, P_cm and q_c, q_m from the three regions
                double Pcc[6], Pmm[6], Pcm[6], qc[3], qm[3];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    Pcm[q] = acc[1][q];
                    Pcc[q] = acc[1][q] + (first_is_c ? acc[0][q] : 0.0) + (last_is_c ? acc[2][q] : 0.0);
                    Pmm[q] = acc[1][q] + (first_is_c ? 0.0 : acc[0][q]) + (last_is_c ? 0.0 : acc[2][q]);
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    qc[q] = acc[1][6 + q] + (first_is_c ? acc[0][6 + q] : 0.0) + (last_is_c ? acc[2][6 + q] : 0.0);
                    qm[q] = acc[1][6 + q] + (first_is_c ? 0.0 : acc[0][6 + q]) + (last_is_c ? 0.0 : acc[2][6 + q]);
                }
                double Amat[36], rhs[6], B[9], t[3], t2[3];
                sym3_sym3(Pcc, lc.Lam, B);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Amat[r * 6 + c] = B[r * 3 + c] + (r == c ? 1.0 : 0.0);
                sym3_sym3(Pcm, lm.Lam, B);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Amat[r * 6 + 3 + c] = B[r * 3 + c];
                sym3_sym3(Pcm, lc.Lam, B);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Amat[(3 + r) * 6 + c] = B[r * 3 + c];
                sym3_sym3(Pmm, lm.Lam, B);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Amat[(3 + r) * 6 + 3 + c] = B[r * 3 + c] + (r == c ? 1.0 : 0.0);
                sym3_mul(Pcc, lc.g, t); sym3_mul(Pcm, lm.g, t2);
#pragma unroll
                for (int q = 0; q < 3; ++q) rhs[q] = qc[q] - t[q] - t2[q];
                sym3_mul(Pcm, lc.g, t); sym3_mul(Pmm, lm.g, t2);
#pragma unroll
                for (int q = 0; q < 3; ++q) rhs[3 + q] = qm[q] - t[q] - t2[q];
                solve_small<6>(Amat, rhs);
                sym3_mul(lc.Lam, rhs, t);
                sym3_mul(lm.Lam, rhs + 3, t2);
#pragma unroll
                for (int q = 0; q < 3; ++q) { zc[q] = lc.g[q] + t[q]; zm[q] = lm.g[q] + t2[q]; }
            }

            // ================= pass B: Gauss-Newton step h_gn = T * prefix(xi) =================
            {
                double tot[3] = {0, 0, 0};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int rb = (r == 0) ? 0 : (r == 1 ? rs : re), rend = (r == 0) ? rs : (r == 1 ? re : L);
                    const int ks = max(k0, rb), ke = min(k1, rend);
                    if (ks >= ke) continue;
                    double z[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (K == 1) z[q] = zc[q];
                        else if (r == 1) z[q] = zc[q] + zm[q];
                        else if (r == 0) z[q] = first_is_c ? zc[q] : zm[q];
                        else z[q] = last_is_c ? zc[q] : zm[q];
                    }
                    Pose2 pa{X[ks], Y[ks], TH[ks]};
                    for (int k = ks; k < ke; ++k) {
                        Pose2 pb{X[k + 1], Y[k + 1], TH[k + 1]};
                        double D[6];
#pragma unroll
                        for (int c = 0; c < 6; ++c) D[c] = DD[(size_t)c * cstride + k];
                        double s, c; sincos(pa.t, &s, &c);
                        double dx = pb.x - pa.x, dy = pb.y - pa.y;
                        double d0 = c * dx + s * dy - ZX[k], d1 = -s * dx + c * dy - ZY[k], d2 = wrap_pi(pb.t - pa.t - ZT[k]);
                        double V[6]; inv_sym3(D, V);
                        const double q02 = pb.y, q12 = -pb.x;
                        // y = Q^T z ; w = V y ; xi = m - Q w = -Q (d + w)
                        double y0 = c * z[0] + s * z[1], y1 = -s * z[0] + c * z[1], y2 = q02 * z[0] + q12 * z[1] + z[2];
                        double w0 = d0 + V[0] * y0 + V[1] * y1 + V[2] * y2;
                        double w1 = d1 + V[1] * y0 + V[3] * y1 + V[4] * y2;
                        double w2 = d2 + V[2] * y0 + V[4] * y1 + V[5] * y2;
                        tot[0] += -(c * w0 - s * w1 + q02 * w2);
                        tot[1] += -(s * w0 + c * w1 + q12 * w2);
                        tot[2] += -w2;
                        HX[k + 1] = tot[0]; HY[k + 1] = tot[1]; HT[k + 1] = tot[2];   // local inclusive prefix
                        pa = pb;
                    }
                }
                block_excl_scan<NT, 3>(tot, red);
                for (int k = k0; k < k1; ++k) {
                    double gx = HX[k + 1] + tot[0], gy = HY[k + 1] + tot[1], gt = HT[k + 1] + tot[2];
                    HX[k + 1] = gx - Y[k + 1] * gt;      // u = T * Xi,  T = [I, S t; 0 1]
                    HY[k + 1] = gy + X[k + 1] * gt;
                    HT[k + 1] = gt;
                }
                bsync<NT>();
            }
            double hh;
            {
                double v[1] = {0};
                for (int k = k0; k < k1; ++k) v[0] += HX[k + 1] * HX[k + 1] + HY[k + 1] * HY[k + 1] + HT[k + 1] * HT[k + 1];
                block_sum<NT, 1>(v, red);
                hh = v[0];
            }
            const double hgnNorm = sqrt(hh);

            // gradient b = -J^T Omega e at vertex j (g2o coordinates), recomputed on demand (rare SD / DL path)
            auto b_vertex = [&](int j, double* bj) {
                bj[0] = bj[1] = bj[2] = 0;
                if (j <= 0) return;
                double gi[3], gj[3];
                {   // edge j-1 -> j, vertex is `to`
                    Pose2 pa{X[j - 1], Y[j - 1], TH[j - 1]}, pb{X[j], Y[j], TH[j]};
                    double D[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) D[c] = DD[(size_t)c * cstride + j - 1];
                    Lin2 e; lin2(pa, pb, ZX[j - 1], ZY[j - 1], ZT[j - 1], D, e);
                    grad2(e, gi, gj);
                    bj[0] -= gj[0]; bj[1] -= gj[1]; bj[2] -= gj[2];
                }
                if (j < L) {   // edge j -> j+1, vertex is `from`
                    Pose2 pa{X[j], Y[j], TH[j]}, pb{X[j + 1], Y[j + 1], TH[j + 1]};
                    double D[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) D[c] = DD[(size_t)c * cstride + j];
                    Lin2 e; lin2(pa, pb, ZX[j], ZY[j], ZT[j], D, e);
                    grad2(e, gi, gj);
                    bj[0] -= gi[0]; bj[1] -= gi[1]; bj[2] -= gi[2];
                }
                grad2(lc.e, gi, gj);
                if (j == lc.jf) { bj[0] -= gi[0]; bj[1] -= gi[1]; bj[2] -= gi[2]; }
                if (j == lc.jt) { bj[0] -= gj[0]; bj[1] -= gj[1]; bj[2] -= gj[2]; }
                if (K == 2) {
                    grad2(lm.e, gi, gj);
                    if (j == lm.jf) { bj[0] -= gi[0]; bj[1] -= gi[1]; bj[2] -= gi[2]; }
                    if (j == lm.jt) { bj[0] -= gj[0]; bj[1] -= gj[1]; bj[2] -= gj[2]; }
                }
            };
            // step at vertex j for the blend hdl = c1 * b + c2 * h_gn
            auto step_vertex = [&](int j, double c1, double c2, double* h) {
                h[0] = c2 * HX[j]; h[1] = c2 * HY[j]; h[2] = c2 * HT[j];
                if (c1 != 0.0) { double bj[3]; b_vertex(j, bj); h[0] += c1 * bj[0]; h[1] += c1 * bj[1]; h[2] += c1 * bj[2]; }
                if (j == 0) { h[0] = h[1] = h[2] = 0; }
            };

            bool sd_ready = false;
            double bb = 0, bHb = 0, bh = 0, alpha = 0, hsdNorm = 0;
            int tries = 0;
            bool good = false;
            const int max_tries = A.max_tries;
            do {
                ++tries;
                double c1 = 0, c2 = 1;       // hdl = c1 * b + c2 * h_gn
                bool is_gn = true;
                if (!(hgnNorm < delta)) {
                    is_gn = false;
                    if (!sd_ready) {
                        // steepest-descent scale: alpha = |b|^2 / (b^T H b), plus b . h_gn
                        double v[3] = {0, 0, 0};
                        double ba[3]; b_vertex(k0, ba);
                        for (int k = k0; k < k1; ++k) {
                            double bbv[3]; b_vertex(k + 1, bbv);
                            Pose2 pa{X[k], Y[k], TH[k]}, pb{X[k + 1], Y[k + 1], TH[k + 1]};
                            double D[6];
#pragma unroll
                            for (int c = 0; c < 6; ++c) D[c] = DD[(size_t)c * cstride + k];
                            Lin2 e; lin2(pa, pb, ZX[k], ZY[k], ZT[k], D, e);
                            double q0, q1, q2; dlin2(e, ba, bbv, q0, q1, q2);
                            v[1] += quad3(D, q0, q1, q2);
                            v[0] += bbv[0] * bbv[0] + bbv[1] * bbv[1] + bbv[2] * bbv[2];
                            v[2] += bbv[0] * HX[k + 1] + bbv[1] * HY[k + 1] + bbv[2] * HT[k + 1];
                            ba[0] = bbv[0]; ba[1] = bbv[1]; ba[2] = bbv[2];
                        }
                        if (tid == 0) {
                            double bf[3], bt[3], q0, q1, q2;
                            b_vertex(lc.jf, bf); b_vertex(lc.jt, bt);
                            dlin2(lc.e, bf, bt, q0, q1, q2); v[1] += quad3(Lc.D, q0, q1, q2);
                            if (K == 2) { b_vertex(lm.jf, bf); b_vertex(lm.jt, bt); dlin2(lm.e, bf, bt, q0, q1, q2); v[1] += quad3(Lm.D, q0, q1, q2); }
                        }
                        block_sum<NT, 3>(v, red);
                        bb = v[0]; bHb = v[1]; bh = v[2];
                        alpha = bb / bHb;
                        hsdNorm = alpha * sqrt(bb);
                        sd_ready = true;
                    }
                    if (hsdNorm > delta) { c1 = delta / hsdNorm * alpha; c2 = 0; }
                    else {
                        const double hsdSq = alpha * alpha * bb;
                        const double c = alpha * bh - hsdSq;                     // hsd . (hgn - hsd)
                        const double bma = hh - 2 * alpha * bh + hsdSq;           // |hgn - hsd|^2
                        double beta;
                        if (c <= 0) beta = (-c + sqrt(c * c + bma * (delta * delta - hsdSq))) / bma;
                        else beta = (delta * delta - hsdSq) / (c + sqrt(c * c + bma * (delta * delta - hsdSq)));
                        c1 = alpha * (1 - beta); c2 = beta;
                    }
                }
                // ================= pass C: try the step =================
                double v[5] = {0, 0, 0, 0, 0};     // newChi, hHh, b.h, |h|^2, (unused)
                double nmax = 0;
                if (k0 < k1) {
                    double ha[3]; step_vertex(k0, c1, c2, ha);
                    Pose2 pa{X[k0], Y[k0], TH[k0]};
                    Pose2 na{pa.x + ha[0], pa.y + ha[1], wrap_pi(pa.t + ha[2])};
                    for (int k = k0; k < k1; ++k) {
                        double hb[3]; step_vertex(k + 1, c1, c2, hb);
                        Pose2 pb{X[k + 1], Y[k + 1], TH[k + 1]};
                        Pose2 nb{pb.x + hb[0], pb.y + hb[1], wrap_pi(pb.t + hb[2])};
                        double D[6];
#pragma unroll
                        for (int c = 0; c < 6; ++c) D[c] = DD[(size_t)c * cstride + k];
                        Lin2 e; lin2(pa, pb, ZX[k], ZY[k], ZT[k], D, e);
                        double q0, q1, q2; dlin2(e, ha, hb, q0, q1, q2);
                        v[1] += quad3(D, q0, q1, q2);
                        v[2] -= e.w0 * q0 + e.w1 * q1 + e.w2 * q2;
                        v[3] += hb[0] * hb[0] + hb[1] * hb[1] + hb[2] * hb[2];
                        Lin2 en; lin2(na, nb, ZX[k], ZY[k], ZT[k], D, en);
                        v[0] += en.chi; nmax = fmax(nmax, en.chi);
                        pa = pb; na = nb; ha[0] = hb[0]; ha[1] = hb[1]; ha[2] = hb[2];
                    }
                }
                double n_cand = 0, n_mem = 0;
                {   // loop edges (every thread computes them: uniform values without another broadcast)
                    double hf[3], ht[3], q0, q1, q2;
                    step_vertex(lc.jf, c1, c2, hf); step_vertex(lc.jt, c1, c2, ht);
                    dlin2(lc.e, hf, ht, q0, q1, q2);
                    Pose2 nf{X[lc.jf] + hf[0], Y[lc.jf] + hf[1], wrap_pi(TH[lc.jf] + hf[2])};
                    Pose2 nt{X[lc.jt] + ht[0], Y[lc.jt] + ht[1], wrap_pi(TH[lc.jt] + ht[2])};
                    Lin2 en; lin2(nf, nt, Lc.meas[0], Lc.meas[1], Lc.meas[2], Lc.D, en);
                    n_cand = en.chi;
                    if (tid == 0) { v[1] += quad3(Lc.D, q0, q1, q2); v[2] -= lc.e.w0 * q0 + lc.e.w1 * q1 + lc.e.w2 * q2; }
                    if (K == 2) {
                        step_vertex(lm.jf, c1, c2, hf); step_vertex(lm.jt, c1, c2, ht);
                        dlin2(lm.e, hf, ht, q0, q1, q2);
                        Pose2 mf{X[lm.jf] + hf[0], Y[lm.jf] + hf[1], wrap_pi(TH[lm.jf] + hf[2])};
                        Pose2 mt{X[lm.jt] + ht[0], Y[lm.jt] + ht[1], wrap_pi(TH[lm.jt] + ht[2])};
                        lin2(mf, mt, Lm.meas[0], Lm.meas[1], Lm.meas[2], Lm.D, en);
                        n_mem = en.chi;
                        if (tid == 0) { v[1] += quad3(Lm.D, q0, q1, q2); v[2] -= lm.e.w0 * q0 + lm.e.w1 * q1 + lm.e.w2 * q2; }
                    }
                }
                block_sum<NT, 5>(v, red);
                nmax = block_max<NT>(nmax, red);
                ++evals;
                const double newChi = v[0] + n_cand + n_mem;
                double linearGain = -v[1] + 2 * v[2];
                const double rawGain = linearGain;
                const double hdlNorm = sqrt(v[3]);
                if (fabs(linearGain) < 1e-12) linearGain = 1e-12;
                const double rho = (cur_chi - newChi) / linearGain;
                if (rho > 0) {
                    good = true;
                    // commit: overwrite H with the step (needs the OLD state for b), then apply
                    if (c1 != 0.0) {
                        double hv[3];
                        // each thread owns vertices k0+1 .. k1
                        double buf_prev[3];
                        (void)buf_prev;
                        // two phases so b_vertex only ever sees the old state
                        for (int k = k0; k < k1; ++k) { step_vertex(k + 1, c1, c2, hv); HX[k + 1] = hv[0]; HY[k + 1] = hv[1]; HT[k + 1] = hv[2]; }
                        bsync<NT>();
                        for (int k = k0; k < k1; ++k) { X[k + 1] += HX[k + 1]; Y[k + 1] += HY[k + 1]; TH[k + 1] = wrap_pi(TH[k + 1] + HT[k + 1]); }
                    } else {
                        for (int k = k0; k < k1; ++k) { X[k + 1] += c2 * HX[k + 1]; Y[k + 1] += c2 * HY[k + 1]; TH[k + 1] = wrap_pi(TH[k + 1] + c2 * HT[k + 1]); }
                    }
                    bsync<NT>();
                    cur_chi = newChi; cur_max = fmax(nmax, fmax(n_cand, n_mem)); cand_chi = n_cand;
                }
                if (rho > 0.75) delta = fmax(delta, 3 * hdlNorm);
                else if (rho < 0.25) delta *= 0.5;
                if (!good) {
                    // a rejected Gauss-Newton step is retried verbatim while it still fits the trust region: every such
                    // retry reproduces the same rho (<= 0), so only the halving of delta and the try counter advance
                    if (is_gn) while (tries < max_tries && hgnNorm < delta) { ++tries; ++evals; delta *= 0.5; }
                    if (A.noise_exit && rawGain <= 1e-9 * cur_chi + 1e-300) { tries = max_tries; }
                }
            } while (!good && tries < max_tries);
            ++iterations;
            if (tries >= max_tries || !good) ok = false;       // Terminate
        }
        if (!have_cur || true) {
            // final computeActiveErrors (src/consensus_utils.cpp:15): the state is the last accepted one and
            // cur_max / cand_chi were taken at exactly that state; recompute only if no iteration ran
            if (!have_cur) {
                loop_lin2(Lc, lo, X, Y, TH, lc, false);
                if (K == 2) loop_lin2(Lm, lo, X, Y, TH, lm, false);
                double v[1] = {0}; double mx = 0;
                if (k0 < k1) {
                    Pose2 pa{X[k0], Y[k0], TH[k0]};
                    for (int k = k0; k < k1; ++k) {
                        Pose2 pb{X[k + 1], Y[k + 1], TH[k + 1]};
                        double D[6];
#pragma unroll
                        for (int c = 0; c < 6; ++c) D[c] = DD[(size_t)c * cstride + k];
                        Lin2 e; lin2(pa, pb, ZX[k], ZY[k], ZT[k], D, e);
                        v[0] += e.chi; mx = fmax(mx, e.chi); pa = pb;
                    }
                }
                block_sum<NT, 1>(v, red);
                mx = block_max<NT>(mx, red);
                cur_chi = v[0] + lc.e.chi + (K == 2 ? lm.e.chi : 0.0);
                cur_max = fmax(mx, fmax(lc.e.chi, K == 2 ? lm.e.chi : 0.0));
                cand_chi = lc.e.chi;
            }
        }
        if (tid == 0) {
            A.verdict[chk] = (cur_max > th) ? 0 : 1;          // every edge chi2 <= th (src/consensus_utils.cpp:17-19)
            if (A.info) {
                ipc_check_info o;
                o.max_chi2 = cur_max; o.cand_chi2 = cand_chi; o.sum_chi2 = cur_chi;
                o.iterations = iterations; o.evals = evals; o.window_len = L; o.n_loops = K;
                A.info[chk] = o;
            }
        }
    }
}

}  // namespace ipcb
