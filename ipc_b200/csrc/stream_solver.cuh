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
    const int* abort;              // host-mapped word (may be null): non-zero = give up at the next iteration (a speculative solve that
                                   // an earlier accept has invalidated)
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

// rank 0: load the diagonal block at (j0, j0), factor it (warp 0, registers), write L_pp back and publish 1 / L_cc (global dinv_g for
// the panel solves of every CTA; shared dinv of the window CTA for the back-substitution). `pre` = the block is already in sm.Dg.
__device__ __forceinline__ void chol_diag_rank0(double* S, int ld, int j0, CholSmem& sm, double* dinv, double* dinv_g, bool pre) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CL_NT / 32;
    if (!pre) {
        for (int c = warp; c < CH_NB; c += NW) sm.Dg[lane][c] = S[(size_t)(j0 + c) * ld + j0 + lane];
    }
    __syncthreads();
    if (warp == 0) chol32_warp(sm, lane);
    __syncthreads();
    for (int c = warp; c < CH_NB; c += NW) if (lane >= c) S[(size_t)(j0 + c) * ld + j0 + lane] = sm.Dg[lane][c];
    if (warp == 0) { dinv[j0 + lane] = sm.inv[lane]; dinv_g[j0 + lane] = sm.inv[lane]; }
}

// X L_pp^T = A_ip for NS row blocks (rb0, rb0 + stride, ...) of one CTA: warp w solves rows w and w + 16 of every block (2 NS
// independent substitution chains), lane c holds a[r][c]. sm.Dg = L_pp, sm.inv = 1 / L_cc.
template <int NS> __device__ __forceinline__ void chol_panel_pass(double* S, int ld, int j0, int rb0, int stride, CholSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CL_NT / 32;
    __syncthreads();
#pragma unroll
    for (int s4 = 0; s4 < NS; ++s4) {
        const int rb = rb0 + s4 * stride;
        for (int c = warp; c < CH_NB; c += NW) sm.Xa[s4][c][lane] = S[(size_t)(j0 + c) * ld + rb * CH_NB + lane];      // Xa[s][c][r]
    }
    __syncthreads();
    double a[2 * NS];
#pragma unroll
    for (int s4 = 0; s4 < NS; ++s4) { a[2 * s4] = sm.Xa[s4][lane][warp]; a[2 * s4 + 1] = sm.Xa[s4][lane][warp + NW]; }
#pragma unroll 4
    for (int c = 0; c < CH_NB; ++c) {
        const double ic = sm.inv[c], lc = sm.Dg[lane][c];
#pragma unroll
        for (int e = 0; e < 2 * NS; ++e) {
            const double x = __shfl_sync(0xffffffffu, a[e], c) * ic;
            a[e] = lane == c ? x : (lane > c ? fma(-x, lc, a[e]) : a[e]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int s4 = 0; s4 < NS; ++s4) { sm.Xa[s4][lane][warp] = a[2 * s4]; sm.Xa[s4][lane][warp + NW] = a[2 * s4 + 1]; }
    __syncthreads();
#pragma unroll
    for (int s4 = 0; s4 < NS; ++s4) {
        const int rb = rb0 + s4 * stride;
        for (int c = warp; c < CH_NB; c += NW) S[(size_t)(j0 + c) * ld + rb * CH_NB + lane] = sm.Xa[s4][c][lane];
    }
}

// dinv: shared memory of the window CTA (null elsewhere); dinv_g: global, n_pad doubles (reuses the force vector z, which is only
// written after the factorisation).
// Schedule per 32-column panel p (look-ahead of one diagonal block):
//   panel  : every CTA that owns row blocks of the panel loads the FACTORED L_pp and solves its rows (up to four 32-row blocks per
//            pass, eight independent substitution chains per warp)                                         -> group barrier
//   update : rank 0 updates the next diagonal block and factors it at once (one warp, registers) while ranks 1.. update the rest of
//            the trailing matrix, one 32 x 32 tile per four warps (rows on lanes, 8 columns per warp as shared-memory broadcasts:
//            bound by the fp64 pipe, not by shared-memory bandwidth); C is fetched together with the X panels  -> group barrier
__device__ __forceinline__ void chol_factor(double* S, int ld, int n_pad, CholSmem& sm, Group& g, double* dinv, double* dinv_g, unsigned long long* prof = nullptr) {
    long long cpt_ = clock64();
    const int nbk = n_pad / CH_NB;             // column blocks
    const int nrb = nbk + 1;                   // row blocks (the last one holds the right-hand side row)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CL_NT / 32;
    const int slot = warp >> 2, q = warp & 3, st = threadIdx.x & 127;
    if (g.rank == 0) chol_diag_rank0(S, ld, 0, sm, dinv, dinv_g, false);
    group_barrier(g);
    CH_PROF(8);
    for (int p = 0; p < nbk; ++p) {
        const int j0 = p * CH_NB;
        // ---- panel solve: row blocks p+1 .. nrb-1 dealt round robin over the group
        const int first_rb = p + 1 + g.rank;
        if (first_rb < nrb) {
            for (int c = warp; c < CH_NB; c += NW) sm.Dg[lane][c] = S[(size_t)(j0 + c) * ld + j0 + lane];
            if (warp == 0) sm.inv[lane] = dinv_g[j0 + lane];
            for (int rb0 = first_rb; rb0 < nrb; rb0 += g.size * CH_SLOTS) {
                const int ns = min(CH_SLOTS, (nrb - rb0 + g.size - 1) / g.size);     // row blocks of this CTA in this pass
                if (ns == 1) chol_panel_pass<1>(S, ld, j0, rb0, g.size, sm);
                else if (ns == 2) chol_panel_pass<2>(S, ld, j0, rb0, g.size, sm);
                else if (ns == 3) chol_panel_pass<3>(S, ld, j0, rb0, g.size, sm);
                else chol_panel_pass<4>(S, ld, j0, rb0, g.size, sm);
            }
        }
        CH_PROF(9);
        group_barrier(g);
        CH_PROF(10);
        // ---- trailing update: C_ij -= X_i X_j^T for block rows i >= j > p
        const int base = p + 1;
        const int m = nbk - base;                                 // remaining column blocks; row blocks: m + 1
        if (m > 0) {
            const int T = (m + 1) * (m + 2) / 2 - 1;              // lower triangle of (m+1) x (m+1) blocks without the last diagonal block
            // tile 0 is the next diagonal block: rank 0 updates it and factors it right away (look-ahead); the other tiles go to
            // ranks 1 .. size-1 (or to everyone but tile 0 when the group is a single CTA)
            const bool solo = g.size == 1;
            const int workers = solo ? 1 : g.size - 1, wrank = solo ? 0 : g.rank - 1;
            if (g.rank == 0) {
                // diagonal tile (base, base): X_i = X_j
                __syncthreads();
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = warp + NW * e;
                    sm.Xa[0][k][lane] = S[(size_t)(j0 + k) * ld + base * CH_NB + lane];
                    sm.Dg[lane][k] = S[(size_t)(base * CH_NB + k) * ld + base * CH_NB + lane];
                }
                __syncthreads();
                {   // thread (row = lane, cols = warp, warp + NW)
                    double acc0 = 0, acc1 = 0;
#pragma unroll 8
                    for (int k = 0; k < CH_NB; ++k) { const double xa = sm.Xa[0][k][lane]; acc0 = fma(xa, sm.Xa[0][k][warp], acc0); acc1 = fma(xa, sm.Xa[0][k][warp + NW], acc1); }
                    sm.Dg[lane][warp] -= acc0; sm.Dg[lane][warp + NW] -= acc1;
                }
                chol_diag_rank0(S, ld, base * CH_NB, sm, dinv, dinv_g, true);
            }
            if (g.rank > 0 || solo) {
                for (int t0 = 1 + wrank; t0 < T; t0 += workers * CH_SLOTS) {
                    const int t = t0 + slot * workers;
                    const bool on = t < T;
                    int ib = 0, jb = 0;
                    if (on) {
                        int I = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                        while (I * (I + 1) / 2 > t) --I;
                        while ((I + 1) * (I + 2) / 2 <= t) ++I;
                        ib = base + I; jb = base + (t - I * (I + 1) / 2);
                    }
                    __syncthreads();
                    double cold[8];
                    double* cptr = S + (size_t)(jb * CH_NB + 8 * q) * ld + ib * CH_NB + lane;
                    if (on) {
                        // slot-local load: 128 threads, X_i and X_j (32 x 32 each); the C values of this thread ride along
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int k = (st >> 5) + 4 * e, r = st & 31;
                            sm.Xa[slot][k][r] = S[(size_t)(j0 + k) * ld + ib * CH_NB + r];
                            sm.Xb[slot][k][r] = S[(size_t)(j0 + k) * ld + jb * CH_NB + r];
                        }
#pragma unroll
                        for (int b8 = 0; b8 < 8; ++b8) cold[b8] = cptr[(size_t)b8 * ld];
                    }
                    __syncthreads();
                    if (on) {
                        double acc[8];
#pragma unroll
                        for (int b8 = 0; b8 < 8; ++b8) acc[b8] = 0;
#pragma unroll 8
                        for (int k = 0; k < CH_NB; ++k) {
                            const double xa = sm.Xa[slot][k][lane];
                            const double4 b0 = *reinterpret_cast<const double4*>(&sm.Xb[slot][k][8 * q]);
                            const double4 b1 = *reinterpret_cast<const double4*>(&sm.Xb[slot][k][8 * q + 4]);
                            acc[0] = fma(xa, b0.x, acc[0]); acc[1] = fma(xa, b0.y, acc[1]); acc[2] = fma(xa, b0.z, acc[2]); acc[3] = fma(xa, b0.w, acc[3]);
                            acc[4] = fma(xa, b1.x, acc[4]); acc[5] = fma(xa, b1.y, acc[5]); acc[6] = fma(xa, b1.z, acc[6]); acc[7] = fma(xa, b1.w, acc[7]);
                        }
#pragma unroll
                        for (int b8 = 0; b8 < 8; ++b8) cptr[(size_t)b8 * ld] = cold[b8] - acc[b8];
                    }
                }
            }
        }
        CH_PROF(11);
        group_barrier(g);
        CH_PROF(12);
    }
}

// Window CTA: L^T z = y with y in matrix row n_pad. zs, dinv: shared memory, n_pad doubles each. Warp w owns columns w and w + 16 of
// a block: both dot products run together, eight loads in flight per lane.
__device__ __forceinline__ void chol_back_substitute(const double* S, int ld, int n_pad, double* zs, const double* dinv, CholSmem& sm, double* z_out) {
    const int nbk = n_pad / CH_NB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CL_NT / 32;
    static_assert(NW * 2 == CH_NB, "two columns per warp");
    for (int b = nbk - 1; b >= 0; --b) {
        const double* c0 = S + (size_t)(b * CH_NB + warp) * ld;
        const double* c1 = S + (size_t)(b * CH_NB + warp + NW) * ld;
        double p0 = 0, p1 = 0, p2 = 0, p3 = 0, r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        int j = (b + 1) * CH_NB + lane;
        for (; j + 96 < n_pad; j += 128) {
            const double a0 = c0[j], a1 = c0[j + 32], a2 = c0[j + 64], a3 = c0[j + 96];
            const double b0 = c1[j], b1 = c1[j + 32], b2 = c1[j + 64], b3 = c1[j + 96];
            const double z0 = zs[j], z1 = zs[j + 32], z2 = zs[j + 64], z3 = zs[j + 96];
            p0 = fma(a0, z0, p0); p1 = fma(a1, z1, p1); p2 = fma(a2, z2, p2); p3 = fma(a3, z3, p3);
            r0 = fma(b0, z0, r0); r1 = fma(b1, z1, r1); r2 = fma(b2, z2, r2); r3 = fma(b3, z3, r3);
        }
        for (; j < n_pad; j += 32) { const double zz = zs[j]; p0 = fma(c0[j], zz, p0); r0 = fma(c1[j], zz, r0); }
        double part0 = (p0 + p1) + (p2 + p3), part1 = (r0 + r1) + (r2 + r3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { part0 += __shfl_xor_sync(0xffffffffu, part0, o); part1 += __shfl_xor_sync(0xffffffffu, part1, o); }
        sm.Dg[lane][warp] = c0[b * CH_NB + lane];                 // diagonal block, Dg[row][col]
        sm.Dg[lane][warp + NW] = c1[b * CH_NB + lane];
        if (lane == 0) { sm.inv[warp] = c0[n_pad] - part0; sm.inv[warp + NW] = c1[n_pad] - part1; }   // y_i - sum_{j beyond the block} L_ji z_j
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
        chol_factor(A.S, A.ld, A.n_pad, cs, g, dinv, A.z, A.prof);
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
                const bool aborted = A.abort && *reinterpret_cast<const volatile int*>(A.abort) != 0;
                A.ctl[0] = (dl.ok && it + 1 < A.max_iter && !aborted) ? 1 : 0;
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
