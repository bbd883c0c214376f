// stream_solver.cuh — the stateful agreementCheck as ONE persistent cooperative kernel per check (sm_100a).
//
// Replaces isAgreeingWithCurrentState (/root/reference/src/consensus_utils.cpp:6-22: initializeOptimization, optimize(iter)
// with g2o's Dogleg, computeActiveErrors, the all-edges chi2 test) for a cluster of ANY size, plus the commit of
// IPC::agreementCheck (/root/reference/src/consensus.cpp:59-71: store / fixComplementary / restore | discard +
// propagateCurrentGuess). Everything runs on the device: one launch, one host synchronisation per check.
//
//   grid  = one CTA of CL_NT threads per SM (cooperative launch: all CTAs are resident, so the hand-rolled grid barrier is safe)
//   CTA 0 = the window: linearisation, prefix scans, gradient, trial states, the Dogleg decisions (cluster_se2/3.cuh)
//   grid  = assembly of the dense force system S (dK x dK, d = 3 | 6) and its blocked right-looking Cholesky:
//           per 32-column panel  [diagonal block factored in shared memory by every CTA that owns rows of the panel]
//           -> [triangular solve of the CTA's 32-row blocks, warp per row, shuffles] -> grid barrier
//           -> [trailing update in 128 x 128 macro tiles, 4 x 4 register tile per thread] -> grid barrier.
//           The right-hand side rides along as one extra matrix row, so the forward substitution comes for free;
//           CTA 0 back-substitutes (coalesced column dot products, one warp per column).
// The Dogleg control flow is that of OptimizationAlgorithmDogleg::solve inside SparseOptimizer::optimize.
#pragma once
#include "cluster_se2.cuh"
#include "cluster_se3.cuh"

namespace ipcb {

constexpr int CH_NB = 32;          // panel width / row-block height
constexpr int CH_MT = 4;           // macro tile = CH_MT x CH_MT blocks (128 x 128)

struct StreamArgs {
    int dim;                       // 2 | 3
    int lo, L, K, n_poses;
    int Lcap;
    const double* odom;            // records the window reads (information x s_factor, or as given for the final optimisation)
    const double* odom_commit;     // records used to re-dead-reckon after an accept (always the s_factor-scaled set)
    double* pose;                  // global vertex estimates of the IPC object
    const void* loops;             // ClLoop | ClLoop3 [K], the candidate last
    ClEvents ev;
    ClBuffers B[2];                // current / trial window state
    double *G, *H;                 // gradient b and h_gn per vertex
    double* lg;                    // per-loop gradient staging
    double* S; int ld, n_pad;      // force system: (n_pad + CH_NB) rows x n_pad columns, column-major; row n_pad = right-hand side
    double* z;                     // forces [n_pad]
    double* res;                   // CL_NRES scalars of the current linearisation / trial (device)
    double* stage3;                // SE(3) dead-reckoning staging
    unsigned* bar;                 // grid barrier counter (zeroed by the host before the launch)
    int* ctl;                      // [0] = 1 while another Gauss-Newton system has to be factorised, 0 = done
    double th; int max_iter; int max_tries; double noise_eps;
    int commit;                    // 1: agreementCheck semantics (store the window on accept + propagateCurrentGuess); 2: final optimisation (always store)
    double* out;                   // results: [0] accepted, [1] max chi2, [2] cand chi2, [3] sum chi2, [4] iterations, [5] evals
};

// ---- grid barrier: monotonic counter, every CTA adds one per barrier; co-residency is guaranteed by the cooperative launch ----
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        const unsigned target = epoch * gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// ---- blocked Cholesky of the lower triangle of S in place; rows beyond n_pad (the right-hand side block) ride along ----------
// smem: Dg[32][33] (diagonal block, becomes L_pp), inv[32], Xa[32][128], Xb[32][128]
struct CholSmem { double Dg[CH_NB][CH_NB + 1]; double inv[CH_NB]; double Xa[CH_NB][CH_MT * CH_NB]; double Xb[CH_NB][CH_MT * CH_NB]; };

__device__ __forceinline__ void chol_factor(double* S, int ld, int n_pad, CholSmem& sm, unsigned* bar, unsigned& epoch) {
    const int nbk = n_pad / CH_NB;             // column blocks
    const int nrb = nbk + 1;                   // row blocks (the last one holds the right-hand side row)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = 0; p < nbk; ++p) {
        const int j0 = p * CH_NB;
        // ---- panel: row blocks p+1 .. nrb-1 are dealt round robin; every owner factors the diagonal block itself
        const int first_rb = p + 1 + (int)blockIdx.x;
        if (first_rb < nrb) {
            // diagonal block -> shared memory (column c by warp c: coalesced)
            sm.Dg[lane][warp] = S[(size_t)(j0 + warp) * ld + j0 + lane];
            __syncthreads();
            // right-looking unblocked Cholesky, thread (i = lane, j = warp), lower triangle i >= j; scaled columns go to Xa[k][i]
            for (int k = 0; k < CH_NB; ++k) {
                const double d = sm.Dg[k][k];
                const double aik = sm.Dg[lane][k], ajk = sm.Dg[warp][k];
                if (warp > k && lane >= warp) sm.Dg[lane][warp] -= aik * ajk / d;
                if (warp == k && lane >= k) sm.Xa[k][lane] = aik * rsqrt(d);
                __syncthreads();
            }
            // L_pp into Dg (lower), 1 / L_cc into inv
            if (lane >= warp) sm.Dg[lane][warp] = sm.Xa[warp][lane];
            __syncthreads();
            if (warp == 0) sm.inv[lane] = 1.0 / sm.Dg[lane][lane];
            if (blockIdx.x == 0 && lane >= warp) S[(size_t)(j0 + warp) * ld + j0 + lane] = sm.Dg[lane][warp];    // CTA 0 always owns a row block
            __syncthreads();
            for (int rb = first_rb; rb < nrb; rb += gridDim.x) {
                const int i0 = rb * CH_NB;
                // X L_pp^T = A_ip, one warp per row r = warp: lane c holds a[r][c]
                // load coalesced through Xb: column c by warp c
                sm.Xb[warp][lane] = S[(size_t)(j0 + warp) * ld + i0 + lane];          // Xb[c][r]
                __syncthreads();
                double a = sm.Xb[lane][warp];                                         // row r = warp, column c = lane
                for (int c = 0; c < CH_NB; ++c) {
                    const double xc = __shfl_sync(0xffffffffu, a, c) * sm.inv[c];
                    if (lane == c) a = xc;
                    else if (lane > c) a -= xc * sm.Dg[lane][c];
                }
                __syncthreads();
                sm.Xb[lane][warp] = a;
                __syncthreads();
                S[(size_t)(j0 + warp) * ld + i0 + lane] = sm.Xb[warp][lane];
                __syncthreads();
            }
        }
        grid_barrier(bar, epoch);
        // ---- trailing update: C_ij -= X_i X_j^T for block rows i >= j > p, macro tiles of CH_MT x CH_MT blocks
        const int base = p + 1;
        const int m_r = nrb - base, m_c = nbk - base;            // remaining row / column blocks
        if (m_c > 0) {
            const int MI = (m_r + CH_MT - 1) / CH_MT, MJ = (m_c + CH_MT - 1) / CH_MT;
            const int ntile = MI * (MI + 1) / 2;
            for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
                int I = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                while (I * (I + 1) / 2 > t) --I;
                while ((I + 1) * (I + 2) / 2 <= t) ++I;
                const int J = t - I * (I + 1) / 2;
                if (J >= MJ) continue;
                const int rb0 = base + I * CH_MT, cb0 = base + J * CH_MT;
                __syncthreads();
                // X panels of the macro row / macro column: Xa[k][r], r < 128
#pragma unroll
                for (int q = 0; q < CH_MT; ++q) {
                    const int rbq = rb0 + q, cbq = cb0 + q;
                    sm.Xa[warp][q * CH_NB + lane] = rbq < nrb ? S[(size_t)(j0 + warp) * ld + rbq * CH_NB + lane] : 0.0;
                    sm.Xb[warp][q * CH_NB + lane] = cbq < nbk ? S[(size_t)(j0 + warp) * ld + cbq * CH_NB + lane] : 0.0;
                }
                __syncthreads();
                double acc[CH_MT][CH_MT];
#pragma unroll
                for (int a = 0; a < CH_MT; ++a)
#pragma unroll
                    for (int b = 0; b < CH_MT; ++b) acc[a][b] = 0;
#pragma unroll 4
                for (int k = 0; k < CH_NB; ++k) {
                    double xa[CH_MT], xb[CH_MT];
#pragma unroll
                    for (int q = 0; q < CH_MT; ++q) { xa[q] = sm.Xa[k][q * CH_NB + lane]; xb[q] = sm.Xb[k][q * CH_NB + warp]; }
#pragma unroll
                    for (int a = 0; a < CH_MT; ++a)
#pragma unroll
                        for (int b = 0; b < CH_MT; ++b) acc[a][b] = fma(xa[a], xb[b], acc[a][b]);
                }
#pragma unroll
                for (int a = 0; a < CH_MT; ++a)
#pragma unroll
                    for (int b = 0; b < CH_MT; ++b) {
                        const int rb = rb0 + a, cb = cb0 + b;
                        if (rb < nrb && cb < nbk && rb >= cb) {
                            double* c = S + (size_t)(cb * CH_NB + warp) * ld + rb * CH_NB + lane;
                            *c -= acc[a][b];
                        }
                    }
            }
        }
        grid_barrier(bar, epoch);
    }
}

// CTA 0: L^T z = y with y in matrix row n_pad. zs: shared memory, n_pad doubles. One warp per column of a block.
__device__ __forceinline__ void chol_back_substitute(const double* S, int ld, int n_pad, double* zs, CholSmem& sm, double* z_out) {
    const int nbk = n_pad / CH_NB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = nbk - 1; b >= 0; --b) {
        const int i = b * CH_NB + warp;                       // this warp's column
        const double* col = S + (size_t)i * ld;
        double part = 0;
        for (int j = (b + 1) * CH_NB + lane; j < n_pad; j += 32) part = fma(col[j], zs[j], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        sm.Dg[lane][warp] = col[b * CH_NB + lane];            // diagonal block, Dg[row][col]
        if (lane == 0) sm.inv[warp] = col[n_pad] - part;      // y_i - sum_{j beyond the block} L_ji z_j
        __syncthreads();
        if (warp == 0) {
            double r = sm.inv[lane];
            for (int c = CH_NB - 1; c >= 0; --c) {
                const double zc = __shfl_sync(0xffffffffu, r, c) / sm.Dg[c][c];
                if (lane == c) r = zc;
                else if (lane < c) r -= sm.Dg[c][lane] * zc;
            }
            zs[b * CH_NB + lane] = r;
            z_out[b * CH_NB + lane] = r;
        }
        __syncthreads();
    }
}

// pad rows / columns of S: identity on the diagonal, zero elsewhere, zero right-hand side block (grid)
__device__ __forceinline__ void chol_init_pad(double* S, int ld, int n, int n_pad) {
    const long long gsz = (long long)gridDim.x * blockDim.x, gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = n_pad + CH_NB;
    // pad columns n .. n_pad-1, all rows
    for (long long e = gid; e < (long long)(n_pad - n) * rows; e += gsz) {
        const int c = n + (int)(e / rows), r = (int)(e % rows);
        S[(size_t)c * ld + r] = (r == c) ? 1.0 : 0.0;
    }
    // real columns: pad rows n .. n_pad-1 and the rows of the right-hand side block other than row n_pad
    const int prow = (n_pad - n) + (CH_NB - 1);
    for (long long e = gid; e < (long long)n * prow; e += gsz) {
        const int c = (int)(e / prow); int r = (int)(e % prow);
        r = r < (n_pad - n) ? n + r : n_pad + 1 + (r - (n_pad - n));
        S[(size_t)c * ld + r] = 0.0;
    }
}

// ---- per-dimension glue -----------------------------------------------------------------------------------------------------
template <int DIM> struct ClDim;
template <> struct ClDim<2> {
    static constexpr int D = 3, PW = 5;
    using Loop = ClLoop;
    static __device__ __forceinline__ void linearize(const StreamArgs& A, int q, double* red) { cl_linearize(A.odom, A.lo, A.L, A.Lcap, A.B[q], A.res, red); }
    static __device__ __forceinline__ void loops(const StreamArgs& A, int q, double* red) { cl_loops(static_cast<const Loop*>(A.loops), A.K, A.B[q], A.res, red); }
    static __device__ __forceinline__ void assemble(const StreamArgs& A, int q) { cl_assemble_grid(static_cast<const Loop*>(A.loops), A.K, A.Lcap, A.B[q], A.S, A.ld, A.n_pad); }
    static __device__ __forceinline__ void gn_step(const StreamArgs& A, int q, double* red) { cl_gn_step(static_cast<const Loop*>(A.loops), A.K, A.ev, A.L, A.Lcap, A.B[q], A.z, A.H, A.res, red); }
    static __device__ __forceinline__ void sd(const StreamArgs& A, int q, double* red) {
        cl_gradient(A.odom, static_cast<const Loop*>(A.loops), A.K, A.ev, A.lo, A.L, A.B[q], A.G, A.lg);
        cl_sd_scalars(A.odom, static_cast<const Loop*>(A.loops), A.K, A.lo, A.L, A.B[q], A.G, A.H, A.res, red);
    }
    static __device__ __forceinline__ void apply(const StreamArgs& A, int q, double c1, double c2, double* red) { cl_apply(A.L, A.B[q].W, A.G, A.H, c1, c2, A.B[q ^ 1].W, A.res, red); }
    static __device__ __forceinline__ void dead_reckon(const StreamArgs& A, int start, double* red) { cl_dead_reckon_cta(A.odom_commit, start, A.n_poses, A.pose, red); }
};
template <> struct ClDim<3> {
    static constexpr int D = 6, PW = 7;
    using Loop = ClLoop3;
    static __device__ __forceinline__ void linearize(const StreamArgs& A, int q, double* red) { cl3_linearize(A.odom, A.lo, A.L, A.Lcap, A.B[q], A.res, red); }
    static __device__ __forceinline__ void loops(const StreamArgs& A, int q, double* red) { cl3_loops(static_cast<const Loop*>(A.loops), A.K, A.B[q], A.res, red); }
    static __device__ __forceinline__ void assemble(const StreamArgs& A, int q) { cl3_assemble_grid(static_cast<const Loop*>(A.loops), A.K, A.Lcap, A.B[q], A.S, A.ld, A.n_pad); }
    static __device__ __forceinline__ void gn_step(const StreamArgs& A, int q, double* red) { cl3_gn_step(static_cast<const Loop*>(A.loops), A.K, A.ev, A.L, A.Lcap, A.B[q], A.z, A.H, A.res, red); }
    static __device__ __forceinline__ void sd(const StreamArgs& A, int q, double* red) {
        cl3_gradient(A.odom, static_cast<const Loop*>(A.loops), A.K, A.ev, A.lo, A.L, A.B[q], A.G, A.lg);
        cl3_sd_scalars(A.odom, static_cast<const Loop*>(A.loops), A.K, A.lo, A.L, A.B[q], A.G, A.H, A.res, red);
    }
    static __device__ __forceinline__ void apply(const StreamArgs& A, int q, double c1, double c2, double* red) { cl3_apply(A.L, A.B[q].W, A.G, A.H, c1, c2, A.B[q ^ 1].W, A.res, red); }
    static __device__ __forceinline__ void dead_reckon(const StreamArgs& A, int start, double* red) { cl3_dead_reckon_cta(A.odom_commit, start, A.n_poses, A.pose, A.stage3); }
};

// Dogleg state of CTA 0 (shared memory; written by thread 0, read by everyone after a barrier)
struct DlState {
    double cur_chi, cur_max, cand_chi, delta;
    double hh, hgnNorm, gn_gain, bb, bh, bHb, alpha, hsdNorm, c1, c2, linearGain;
    int cur, iterations, evals, tries, ok, have_sd, good, trial_gn, go_on, need_sd;
};

template <int DIM>
__global__ void __launch_bounds__(CL_NT, 1) stream_check_kernel(StreamArgs A) {
    using T = ClDim<DIM>;
    extern __shared__ __align__(16) unsigned char smraw[];
    CholSmem& cs = *reinterpret_cast<CholSmem*>(smraw);
    double* red = reinterpret_cast<double*>(smraw + sizeof(CholSmem));                 // 32 x 27 staging
    DlState& dl = *reinterpret_cast<DlState*>(red + 32 * 27);
    double* zs = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(&dl) + ((sizeof(DlState) + 15) & ~15));   // n_pad doubles
    unsigned epoch = 0;
    const bool cta0 = blockIdx.x == 0;
    const int n = T::D * A.K;
    double* res = A.res;

    chol_init_pad(A.S, A.ld, n, A.n_pad);
    if (cta0) {
        cl_copy_cta(A.pose + (size_t)T::PW * A.lo, A.B[0].W, (long long)T::PW * (A.L + 1));
        T::linearize(A, 0, red);
        T::loops(A, 0, red);
        if (threadIdx.x == 0) {
            dl.cur = 0; dl.cur_chi = res[0] + res[2]; dl.cur_max = fmax(res[1], res[3]); dl.cand_chi = res[4];
            dl.delta = 1e4; dl.iterations = 0; dl.evals = 0; dl.ok = 1;
            A.ctl[0] = A.max_iter > 0 ? 1 : 0; A.ctl[1] = 0;
        }
        __syncthreads();
    }
    grid_barrier(A.bar, epoch);
    for (int it = 0; it < A.max_iter; ++it) {
        if (*reinterpret_cast<volatile int*>(A.ctl) == 0) break;
        // ---- Gauss-Newton system of the current linearisation: assemble (grid), factor (grid), back-substitute (CTA 0)
        const int cur = cta0 ? dl.cur : *reinterpret_cast<volatile int*>(A.ctl + 1);
        T::assemble(A, cur);
        grid_barrier(A.bar, epoch);
        chol_factor(A.S, A.ld, A.n_pad, cs, A.bar, epoch);
        if (cta0) {
            chol_back_substitute(A.S, A.ld, A.n_pad, zs, cs, A.z);
            __threadfence_block();
            __syncthreads();
            T::gn_step(A, cur, red);
            if (threadIdx.x == 0) {
                dl.hh = res[5]; dl.hgnNorm = sqrt(res[5]); dl.gn_gain = res[6];
                dl.have_sd = 0; dl.good = 0; dl.tries = 0;
                dl.go_on = isfinite(dl.hgnNorm) ? 1 : 0;           // factorisation broke down (g2o: Fail)
                if (!dl.go_on) { dl.ok = 0; ++dl.iterations; }
            }
            __syncthreads();
            while (dl.go_on) {
                if (threadIdx.x == 0) {
                    ++dl.tries;
                    dl.c1 = 0; dl.c2 = 1; dl.linearGain = dl.gn_gain;
                    dl.trial_gn = dl.hgnNorm < dl.delta;
                    dl.need_sd = !dl.trial_gn && !dl.have_sd;
                }
                __syncthreads();
                if (dl.need_sd) {
                    T::sd(A, dl.cur, red);
                    if (threadIdx.x == 0) {
                        dl.bb = res[7]; dl.bh = res[8]; dl.bHb = res[9];
                        dl.alpha = dl.bb / dl.bHb; dl.hsdNorm = dl.alpha * sqrt(dl.bb); dl.have_sd = 1;
                    }
                    __syncthreads();
                }
                if (threadIdx.x == 0 && !dl.trial_gn) {
                    const double delta = dl.delta, alpha = dl.alpha, bb = dl.bb, bh = dl.bh, bHb = dl.bHb, hh = dl.hh;
                    double c1, c2;
                    if (dl.hsdNorm > delta) { c1 = delta / dl.hsdNorm * alpha; c2 = 0; }
                    else {
                        const double hsdSq = alpha * alpha * bb;
                        const double c = alpha * bh - hsdSq, bma = hh - 2 * alpha * bh + hsdSq;
                        double beta;
                        if (c <= 0) beta = (-c + sqrt(c * c + bma * (delta * delta - hsdSq))) / bma;
                        else beta = (delta * delta - hsdSq) / (c + sqrt(c * c + bma * (delta * delta - hsdSq)));
                        c1 = alpha * (1 - beta); c2 = beta;
                    }
                    dl.c1 = c1; dl.c2 = c2;
                    dl.linearGain = -(c1 * c1 * bHb + 2 * c1 * c2 * bb + c2 * c2 * bh) + 2 * (c1 * bb + c2 * bh);
                }
                __syncthreads();
                const int q = dl.cur;
                T::apply(A, q, dl.c1, dl.c2, red);
                T::linearize(A, q ^ 1, red);
                T::loops(A, q ^ 1, red);
                if (threadIdx.x == 0) {
                    ++dl.evals;
                    const double newChi = res[0] + res[2], hdlNorm = sqrt(res[10]);
                    double linearGain = dl.linearGain;
                    const double rawGain = linearGain;
                    if (fabs(linearGain) < 1e-12) linearGain = 1e-12;
                    const double rho = (dl.cur_chi - newChi) / linearGain;
                    if (rho > 0) { dl.good = 1; dl.cur = q ^ 1; dl.cur_chi = newChi; dl.cur_max = fmax(res[1], res[3]); dl.cand_chi = res[4]; }
                    if (rho > 0.75) dl.delta = fmax(dl.delta, 3 * hdlNorm);
                    else if (rho < 0.25) dl.delta *= 0.5;
                    if (!dl.good) {
                        // a rejected Gauss-Newton step is retried verbatim while it still fits the trust region: same rho, only delta halves
                        if (dl.trial_gn) while (dl.tries < A.max_tries && dl.hgnNorm < dl.delta) { ++dl.tries; ++dl.evals; dl.delta *= 0.5; }
                        if (A.noise_eps > 0 && rawGain <= A.noise_eps * dl.cur_chi + 1e-300) dl.tries = A.max_tries;
                    }
                    dl.go_on = (!dl.good && dl.tries < A.max_tries) ? 1 : 0;
                    if (!dl.go_on) {
                        ++dl.iterations;
                        if (dl.tries >= A.max_tries || !dl.good) dl.ok = 0;
                    }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                A.ctl[1] = dl.cur;
                A.ctl[0] = (dl.ok && it + 1 < A.max_iter) ? 1 : 0;
                __threadfence();
            }
            __syncthreads();
        }
        grid_barrier(A.bar, epoch);
    }
    if (cta0) {
        const bool accepted = !(dl.cur_max > A.th);
        if (threadIdx.x == 0) {
            A.out[0] = accepted ? 1.0 : 0.0; A.out[1] = dl.cur_max; A.out[2] = dl.cand_chi; A.out[3] = dl.cur_chi;
            A.out[4] = dl.iterations; A.out[5] = dl.evals;
        }
        if ((A.commit == 1 && accepted) || A.commit == 2) {
            // discard + propagateCurrentGuess (src/consensus.cpp:69-71); a rejection leaves the global estimates untouched (restore)
            cl_copy_cta(A.B[dl.cur].W, A.pose + (size_t)T::PW * A.lo, (long long)T::PW * (A.L + 1));
            if (A.commit == 1) T::dead_reckon(A, A.lo + A.L, red);
        }
    }
}

inline size_t stream_smem_bytes(int n_pad) { return sizeof(CholSmem) + sizeof(double) * 32 * 27 + ((sizeof(DlState) + 15) & ~15) + sizeof(double) * (size_t)n_pad + 64; }

}  // namespace ipcb
