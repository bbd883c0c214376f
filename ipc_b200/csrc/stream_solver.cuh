// stream_solver.cuh — the stateful agreementCheck as ONE persistent cooperative kernel per check (sm_100a).
//
// Replaces isAgreeingWithCurrentState (/root/reference/src/consensus_utils.cpp:6-22: initializeOptimization, optimize(iter)
// with g2o's Dogleg, computeActiveErrors, the all-edges chi2 test) for a cluster of ANY size, plus the commit of
// IPC::agreementCheck (/root/reference/src/consensus.cpp:59-71: store / fixComplementary / restore | discard +
// propagateCurrentGuess). Everything runs on the device: one launch, one host synchronisation per check.
//
//   grid   = one CTA of CL_NT threads per SM (cooperative launch: all CTAs are resident, so the hand-rolled barriers are safe)
//   group  = the CTAs that solve ONE check: the whole grid, or a slice of it when the stream speculates on several candidates
//   rank 0 = the window CTA: linearisation, prefix scans, gradient, trial states, the Dogleg decisions (cluster_se2/3.cuh)
//   group  = assembly of the dense force system S (dK x dK, d = 3 | 6) and its blocked right-looking Cholesky:
//            per 32-column panel  [diagonal block factored by one warp in registers in every CTA that owns rows of the panel]
//            -> [triangular solve of the CTA's 32-row blocks, two rows per warp, shuffles] -> group barrier
//            -> [trailing update, one 32 x 32 tile per four warps, 1 x 8 register tile per thread] -> group barrier.
//            The right-hand side rides along as one extra matrix row, so the forward substitution comes for free;
//            the window CTA back-substitutes (coalesced column dot products, one warp per column).
// The Dogleg control flow is that of OptimizationAlgorithmDogleg::solve inside SparseOptimizer::optimize.
#pragma once
#include "cluster_se2.cuh"
#include "cluster_se3.cuh"

namespace ipcb {

constexpr int CH_NB = 32;          // panel width / row-block height
constexpr int CH_SLOTS = CL_NT / 128;   // trailing-update tile slots per CTA: four warps per 32 x 32 tile

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
    unsigned* bar;                 // group barrier counter (zeroed by the host before the launch)
    int* ctl;                      // [0] = 1 while another Gauss-Newton system has to be factorised, 0 = done; [1] = current state buffer
    double th; int max_iter; int max_tries; double noise_eps;
    int commit;                    // 1: agreementCheck semantics (store the window on accept + propagateCurrentGuess); 2: final optimisation
                                   // (always store); 0: leave the global estimates alone (speculative slot: the host commits in order)
    double* out;                   // results: [0] accepted, [1] max chi2, [2] cand chi2, [3] sum chi2, [4] iterations, [5] evals,
                                   // [6] factorisations, [7] trial states evaluated, [8] index of the final state buffer
    unsigned long long* prof;      // window CTA / thread 0 cycle counters per phase (ipc_stream_profile): 0 setup, 1 assemble, 2 factor,
                                   // 3 back-substitution, 4 GN step, 5 steepest descent, 6 trial states, 7 commit; 8.. factorisation parts
};
#define ST_PROF(i) do { if (cta0 && threadIdx.x == 0 && A.prof) { const long long now_ = clock64(); A.prof[i] += (unsigned long long)(now_ - pt_); pt_ = now_; } } while (0)

// A check is solved by a GROUP of consecutive CTAs of the cooperative grid (all CTAs are co-resident, so spinning is safe): the
// whole grid for one check, or several groups side by side when the stream speculates on several candidates at once.
struct Group { int rank, size; unsigned* bar; unsigned epoch; };

// ---- group barrier: monotonic counter, every CTA of the group adds one per barrier ----
__device__ __forceinline__ void group_barrier(Group& g) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++g.epoch;
        const unsigned target = g.epoch * (unsigned)g.size;
        __threadfence();
        atomicAdd(g.bar, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(g.bar) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// ---- blocked Cholesky of the lower triangle of S in place; rows beyond n_pad (the right-hand side block) ride along ----------
struct CholSmem {
    double Dg[CH_NB][CH_NB + 1];                 // diagonal block, becomes L_pp
    double inv[CH_NB];                           // 1 / L_cc
    double Xa[CH_SLOTS][CH_NB][CH_NB];           // per tile slot: X_i as [k][r]   (also the row-block staging of the panel solve)
    double Xb[CH_SLOTS][CH_NB][CH_NB];           // per tile slot: X_j as [k][c]
};

// 32 x 32 Cholesky by ONE warp, row `lane` in registers, no block barrier: per pivot one broadcast, one rsqrt, one rank-1 update
__device__ __noinline__ void chol32_warp(CholSmem& sm, int lane) {
    double a[CH_NB];
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) a[c] = sm.Dg[lane][c];
#pragma unroll
    for (int k = 0; k < CH_NB; ++k) {
        const double d = __shfl_sync(0xffffffffu, a[k], k);
        const double rs = rsqrt(d);
        const double l = a[k] * rs;                 // L[lane][k] for lane >= k
        a[k] = l;
        if (lane == k) sm.inv[k] = rs;
#pragma unroll
        for (int j = k + 1; j < CH_NB; ++j) a[j] = fma(-l, __shfl_sync(0xffffffffu, l, j), a[j]);
    }
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) if (lane >= c) sm.Dg[lane][c] = a[c];
}

#define CH_PROF(i) do { if (prof && g.rank == 0 && threadIdx.x == 0) { const long long now_ = clock64(); prof[i] += (unsigned long long)(now_ - cpt_); cpt_ = now_; } } while (0)
// dinv (shared memory of the window CTA, may be null elsewhere): 1 / L_cc of every column, for the back-substitution
__device__ __forceinline__ void chol_factor(double* S, int ld, int n_pad, CholSmem& sm, Group& g, double* dinv, unsigned long long* prof = nullptr) {
    long long cpt_ = clock64();
    const int nbk = n_pad / CH_NB;             // column blocks
    const int nrb = nbk + 1;                   // row blocks (the last one holds the right-hand side row)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CL_NT / 32;
    for (int p = 0; p < nbk; ++p) {
        const int j0 = p * CH_NB;
        // ---- panel: row blocks p+1 .. nrb-1 are dealt round robin; every owner factors the diagonal block itself
        const int first_rb = p + 1 + g.rank;
        if (first_rb < nrb) {
            // diagonal block and this CTA's first row block -> shared memory (column c by warp: coalesced)
            for (int c = warp; c < CH_NB; c += NW) {
                sm.Dg[lane][c] = S[(size_t)(j0 + c) * ld + j0 + lane];
                sm.Xa[0][c][lane] = S[(size_t)(j0 + c) * ld + first_rb * CH_NB + lane];      // [c][r]
            }
            __syncthreads();
            if (warp == 0) chol32_warp(sm, lane);
            __syncthreads();
            if (g.rank == 0) {                  // the window CTA always owns a row block: it writes L_pp back and keeps 1 / L_cc
                for (int c = warp; c < CH_NB; c += NW) if (lane >= c) S[(size_t)(j0 + c) * ld + j0 + lane] = sm.Dg[lane][c];
                if (dinv && warp == 0) dinv[j0 + lane] = sm.inv[lane];
            }
            CH_PROF(8);
            for (int rb = first_rb; rb < nrb; rb += g.size) {
                const int i0 = rb * CH_NB;
                if (rb != first_rb) {
                    __syncthreads();
                    for (int c = warp; c < CH_NB; c += NW) sm.Xa[0][c][lane] = S[(size_t)(j0 + c) * ld + i0 + lane];
                    __syncthreads();
                }
                // X L_pp^T = A_ip: warp w solves rows w and w + NW (two independent chains), lane c holds a[r][c]
                double a0 = sm.Xa[0][lane][warp], a1 = sm.Xa[0][lane][warp + NW];
#pragma unroll 4
                for (int c = 0; c < CH_NB; ++c) {
                    const double ic = sm.inv[c], lc = sm.Dg[lane][c];
                    const double x0 = __shfl_sync(0xffffffffu, a0, c) * ic, x1 = __shfl_sync(0xffffffffu, a1, c) * ic;
                    if (lane == c) { a0 = x0; a1 = x1; }
                    else if (lane > c) { a0 = fma(-x0, lc, a0); a1 = fma(-x1, lc, a1); }
                }
                __syncthreads();
                sm.Xa[0][lane][warp] = a0; sm.Xa[0][lane][warp + NW] = a1;
                __syncthreads();
                for (int c = warp; c < CH_NB; c += NW) S[(size_t)(j0 + c) * ld + i0 + lane] = sm.Xa[0][c][lane];
            }
            CH_PROF(9);
        }
        group_barrier(g);
        CH_PROF(10);
        // ---- trailing update: C_ij -= X_i X_j^T for block rows i >= j > p. One 32 x 32 tile per slot of four warps (warp q of the
        // slot owns columns 8q .. 8q+7 of the tile: rows come from distinct lanes, columns are shared-memory broadcasts, so the loop
        // is bound by the fp64 pipe, not by shared-memory bandwidth). Tiles are dealt to CTAs first, then to slots.
        const int base = p + 1;
        const int m = nbk - base;                                 // remaining column blocks; row blocks: m + 1
        if (m > 0) {
            const int T = (m + 1) * (m + 2) / 2 - 1;              // lower triangle of (m+1) x (m+1) blocks without the last diagonal block
            const int slot = warp >> 2, q = warp & 3, st = threadIdx.x & 127;
            for (int t0 = g.rank; t0 < T; t0 += g.size * CH_SLOTS) {
                const int t = t0 + slot * g.size;
                const bool on = t < T;
                int ib = 0, jb = 0;
                if (on) {
                    int I = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                    while (I * (I + 1) / 2 > t) --I;
                    while ((I + 1) * (I + 2) / 2 <= t) ++I;
                    ib = base + I; jb = base + (t - I * (I + 1) / 2);
                }
                __syncthreads();
                if (on) {
                    // slot-local load: 128 threads, X_i and X_j (32 x 32 each): thread st loads rows (st & 31), columns (st >> 5) + 4 e
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = (st >> 5) + 4 * e, r = st & 31;
                        sm.Xa[slot][k][r] = S[(size_t)(j0 + k) * ld + ib * CH_NB + r];
                        sm.Xb[slot][k][r] = S[(size_t)(j0 + k) * ld + jb * CH_NB + r];
                    }
                }
                __syncthreads();
                if (on) {
                    double acc[8];
#pragma unroll
                    for (int b = 0; b < 8; ++b) acc[b] = 0;
#pragma unroll 8
                    for (int k = 0; k < CH_NB; ++k) {
                        const double xa = sm.Xa[slot][k][lane];
                        const double4 b0 = *reinterpret_cast<const double4*>(&sm.Xb[slot][k][8 * q]);
                        const double4 b1 = *reinterpret_cast<const double4*>(&sm.Xb[slot][k][8 * q + 4]);
                        acc[0] = fma(xa, b0.x, acc[0]); acc[1] = fma(xa, b0.y, acc[1]); acc[2] = fma(xa, b0.z, acc[2]); acc[3] = fma(xa, b0.w, acc[3]);
                        acc[4] = fma(xa, b1.x, acc[4]); acc[5] = fma(xa, b1.y, acc[5]); acc[6] = fma(xa, b1.z, acc[6]); acc[7] = fma(xa, b1.w, acc[7]);
                    }
                    double* c = S + (size_t)(jb * CH_NB + 8 * q) * ld + ib * CH_NB + lane;
#pragma unroll
                    for (int b = 0; b < 8; ++b) c[(size_t)b * ld] -= acc[b];
                }
            }
        }
        CH_PROF(11);
        group_barrier(g);
        CH_PROF(12);
    }
}

// Window CTA: L^T z = y with y in matrix row n_pad. zs, dinv: shared memory, n_pad doubles each. One warp per column of a block.
__device__ __forceinline__ void chol_back_substitute(const double* S, int ld, int n_pad, double* zs, const double* dinv, CholSmem& sm, double* z_out) {
    const int nbk = n_pad / CH_NB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CL_NT / 32;
    for (int b = nbk - 1; b >= 0; --b) {
        for (int w = warp; w < CH_NB; w += NW) {
            const int i = b * CH_NB + w;                          // this warp's column
            const double* col = S + (size_t)i * ld;
            double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
            int j = (b + 1) * CH_NB + lane;
            for (; j + 96 < n_pad; j += 128) {                    // four loads in flight per lane
                const double c0 = col[j], c1 = col[j + 32], c2 = col[j + 64], c3 = col[j + 96];
                p0 = fma(c0, zs[j], p0); p1 = fma(c1, zs[j + 32], p1); p2 = fma(c2, zs[j + 64], p2); p3 = fma(c3, zs[j + 96], p3);
            }
            for (; j < n_pad; j += 32) p0 = fma(col[j], zs[j], p0);
            double part = (p0 + p1) + (p2 + p3);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            sm.Dg[lane][w] = col[b * CH_NB + lane];               // diagonal block, Dg[row][col]
            if (lane == 0) sm.inv[w] = col[n_pad] - part;         // y_i - sum_{j beyond the block} L_ji z_j
        }
        __syncthreads();
        if (warp == 0) {
            double r = sm.inv[lane];
            const double di = dinv[b * CH_NB + lane];
#pragma unroll 4
            for (int c = CH_NB - 1; c >= 0; --c) {
                const double zc = __shfl_sync(0xffffffffu, r * di, c);
                if (lane == c) r = zc;
                else if (lane < c) r = fma(-sm.Dg[c][lane], zc, r);
            }
            zs[b * CH_NB + lane] = r;
            z_out[b * CH_NB + lane] = r;
        }
        __syncthreads();
    }
}

// pad rows / columns of S: identity on the diagonal, zero elsewhere, zero right-hand side block (group)
__device__ __forceinline__ void chol_init_pad(double* S, int ld, int n, int n_pad, const Group& g) {
    const long long gsz = (long long)g.size * blockDim.x, gid = (long long)g.rank * blockDim.x + threadIdx.x;
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
    static __device__ __forceinline__ void assemble(const StreamArgs& A, int q, const Group& g) { cl_assemble_grid(static_cast<const Loop*>(A.loops), A.K, A.Lcap, A.B[q], A.S, A.ld, A.n_pad, g.rank, g.size); }
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
    static __device__ __forceinline__ void assemble(const StreamArgs& A, int q, const Group& g) { cl3_assemble_grid(static_cast<const Loop*>(A.loops), A.K, A.Lcap, A.B[q], A.S, A.ld, A.n_pad, g.rank, g.size); }
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
__global__ void __launch_bounds__(CL_NT, 1) stream_check_kernel(const StreamArgs* __restrict__ all_args, int group_size) {
    using T = ClDim<DIM>;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ StreamArgs A;
    {
        const int* src = reinterpret_cast<const int*>(all_args + blockIdx.x / group_size);
        for (int i = threadIdx.x; i < (int)(sizeof(StreamArgs) / sizeof(int)); i += blockDim.x) reinterpret_cast<int*>(&A)[i] = src[i];
    }
    __syncthreads();
    if (A.K <= 0) return;                                  // idle speculation slot
    Group g{(int)(blockIdx.x % group_size), group_size, A.bar, 0u};
    CholSmem& cs = *reinterpret_cast<CholSmem*>(smraw);
    double* red = reinterpret_cast<double*>(smraw + sizeof(CholSmem));                 // 32 x 27 staging
    DlState& dl = *reinterpret_cast<DlState*>(red + 32 * 27);
    double* zs = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(&dl) + ((sizeof(DlState) + 15) & ~15));   // n_pad doubles
    double* dinv = zs + A.n_pad;                                                                                      // n_pad doubles
    const bool cta0 = g.rank == 0;
    const int n = T::D * A.K;
    double* res = A.res;
    long long pt_ = clock64();
    int n_fact = 0, n_trial = 0;

    chol_init_pad(A.S, A.ld, n, A.n_pad, g);
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
    group_barrier(g);
    ST_PROF(0);
    for (int it = 0; it < A.max_iter; ++it) {
        if (*reinterpret_cast<volatile int*>(A.ctl) == 0) break;
        // ---- Gauss-Newton system of the current linearisation: assemble (grid), factor (grid), back-substitute (CTA 0)
        const int cur = cta0 ? dl.cur : *reinterpret_cast<volatile int*>(A.ctl + 1);
        T::assemble(A, cur, g);
        group_barrier(g);
        ST_PROF(1);
        chol_factor(A.S, A.ld, A.n_pad, cs, g, cta0 ? dinv : nullptr, A.prof);
        ST_PROF(2);
        ++n_fact;
        if (cta0) {
            chol_back_substitute(A.S, A.ld, A.n_pad, zs, dinv, cs, A.z);
            __threadfence_block();
            __syncthreads();
            ST_PROF(3);
            T::gn_step(A, cur, red);
            ST_PROF(4);
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
                    ST_PROF(6);
                    T::sd(A, dl.cur, red);
                    ST_PROF(5);
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
                ++n_trial;
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
            ST_PROF(6);
            if (threadIdx.x == 0) {
                A.ctl[1] = dl.cur;
                A.ctl[0] = (dl.ok && it + 1 < A.max_iter) ? 1 : 0;
                __threadfence();
            }
            __syncthreads();
        }
        group_barrier(g);
    }
    if (cta0) {
        const bool accepted = !(dl.cur_max > A.th);
        if (threadIdx.x == 0) {
            A.out[0] = accepted ? 1.0 : 0.0; A.out[1] = dl.cur_max; A.out[2] = dl.cand_chi; A.out[3] = dl.cur_chi;
            A.out[4] = dl.iterations; A.out[5] = dl.evals; A.out[6] = n_fact; A.out[7] = n_trial; A.out[8] = dl.cur;
        }
        if ((A.commit == 1 && accepted) || A.commit == 2) {
            // discard + propagateCurrentGuess (src/consensus.cpp:69-71); a rejection leaves the global estimates untouched (restore)
            cl_copy_cta(A.B[dl.cur].W, A.pose + (size_t)T::PW * A.lo, (long long)T::PW * (A.L + 1));
            if (A.commit == 1) T::dead_reckon(A, A.lo + A.L, red);
        }
        ST_PROF(7);
    }
}

// commit of a speculative slot: store its window, re-dead-reckon everything after it (src/consensus.cpp:69-71)
template <int DIM>
__global__ void __launch_bounds__(CL_NT) stream_commit_kernel(const double* W, double* pose, int lo, int L, const double* odom, int n_poses, double* stage3) {
    __shared__ double red[32 * 2];
    cl_copy_cta(W, pose + (size_t)ClDim<DIM>::PW * lo, (long long)ClDim<DIM>::PW * (L + 1));
    if (DIM == 2) cl_dead_reckon_cta(odom, lo + L, n_poses, pose, red);
    else cl3_dead_reckon_cta(odom, lo + L, n_poses, pose, stage3);
}

inline size_t stream_smem_bytes(int n_pad) { return sizeof(CholSmem) + sizeof(double) * 32 * 27 + ((sizeof(DlState) + 15) & ~15) + sizeof(double) * 2 * (size_t)n_pad + 64; }

}  // namespace ipcb
