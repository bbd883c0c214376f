// chain_se2.cuh — SE(2) window check, one CTA per check (K = 1 or 2 loop edges), ONE pass per Dogleg iteration.
//
// Replaces, for independent fast-path / pair sub-problems, the reference's
//   isAgreeingWithCurrentState  /root/reference/src/consensus_utils.cpp:6-22
// driven as in IPC::agreementCheck /root/reference/src/consensus.cpp:42-75 from the dead-reckoned state of
// IPC::IPC (:9-33), with g2o's Dogleg semantics.
//
// Formulation (DESIGN.md "Chain solve"). With the twist of odometry edge k as unknown,
//   xi_k = Q_k q_k,  Q_k = [R_k, -S t_{k+1}; 0 1],   u_j = T_j Xi_j,  Xi_j = sum_{k<j} xi_k,  T_j = [I, S t_j; 0 1]
// the odometry part of the Gauss-Newton matrix is block diagonal and every loop edge is a rank-3 term on an
// interval of edges, so the exact GN step is
//   Xi_j = Pm_j - PM_j z_rho(j) - C_rho(j),     PM_j = sum_{k<j} Q_k V_k Q_k^T,  Pm_j = -sum_{k<j} Q_k d_k
// where the per-region "forces" z come from a 3x3 (K = 1) or 6x6 (K = 2) capacitance solve on interval sums of PM, Pm.
// One sweep over the window therefore (i) re-derives the prefix sums of the OLD linearisation on the fly (cheap: the
// cos / sin of every pose are part of the state), (ii) applies the step to every pose, (iii) re-linearises every edge at the
// new pose (one sincos per pose), (iv) accumulates chi2 / max chi2 of the trial state and (v) produces the interval sums
// of PM, Pm of the NEW linearisation, i.e. everything the next capacitance solve needs. State: 5 doubles per vertex.
// The predicted gain of a GN step is chi2 - model(h_gn) with model = sum_r z_r^T P_r z_r + loop terms (O(1)).
// Trust-region-limited steps (steepest-descent / dog-leg blends) need the gradient in g2o's vertex coordinates:
// two extra sweeps (b, b^T H b) into a per-CTA scratch, then the same trial sweep reading h = c1 b + c2 h_gn.
// Rejected trials roll the poses back from a per-CTA backup (written by the sweep) and re-linearise.
//
// The sweep code is __host__ __device__: tests/ compiles it with NT = 1 on the CPU to validate the arithmetic and
// the control flow against the oracle without a GPU (tests/host_emul/); the product only ever runs the CUDA build.
#pragma once
#include "common.cuh"

namespace ipcb {

#ifdef __CUDACC__
#define IPC_HD __host__ __device__ __forceinline__
#define IPC_HD_COLD inline __host__ __device__ __noinline__      // cold / once-per-sweep code: keep its registers out of the hot loop
#else
#define IPC_HD inline
#define IPC_HD_COLD inline
#endif

// sin / cos for |t| <= pi + small (every angle here is normalised): Cody-Waite reduction by multiples of pi/2 and the
// fdlibm kernel polynomials on [-pi/4, pi/4] (< 1 ulp each). The CUDA library sincos carries a Payne-Hanek slow path
// (local-memory table) that this kernel never needs. Coefficients sit in constant memory so the FMAs read them as operands.
#ifdef __CUDACC__
static __constant__ double kSinCos[12] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                   -1.98412698298579493134e-04, 8.33333333332248946124e-03,  -1.66666666666666324348e-01,
                                   -1.13596475577881948265e-11, 2.08757232129817482790e-09,  -2.75573143513906633035e-07,
                                   2.48015872894767294178e-05,  -1.38888888888741095749e-03, 4.16666666666666019037e-02};
#endif
IPC_HD void ipc_sincos(double t, double* s, double* c) {
#ifdef __CUDA_ARCH__
    const double q = rint(t * 0.63661977236758134308);            // 2 / pi
    double r = fma(-q, 1.57079632679489655800e+00, t);             // pi/2 hi
    r = fma(-q, 6.12323399573676603587e-17, r);                    // pi/2 lo
    const double z = r * r;
    double ps = fma(z, kSinCos[0], kSinCos[1]);
    ps = fma(z, ps, kSinCos[2]);
    ps = fma(z, ps, kSinCos[3]);
    ps = fma(z, ps, kSinCos[4]);
    ps = fma(z, ps, kSinCos[5]);
    const double sr = fma(z * r, ps, r);
    double pc = fma(z, kSinCos[6], kSinCos[7]);
    pc = fma(z, pc, kSinCos[8]);
    pc = fma(z, pc, kSinCos[9]);
    pc = fma(z, pc, kSinCos[10]);
    pc = fma(z, pc, kSinCos[11]);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int n = (int)q;
    const double ss = (n & 1) ? cr : sr, cc = (n & 1) ? sr : cr;
    *s = (n & 2) ? -ss : ss;
    *c = ((n + 1) & 2) ? -cc : cc;
#else
    *s = sin(t); *c = cos(t);
#endif
}
// g2o normalize_theta maps to [-pi, pi). Branch-free form: subtract the nearest multiple of 2 pi (exact no-op for the
// common in-range case; the two differ only for |t| = pi exactly, where the residual / heading is equivalent).
IPC_HD double wrap_pi_hd(double t) {
    const double k = rint(t * 0.15915494309189533577);             // 1 / (2 pi)
    return fma(-k, 6.28318530717958647692, t);
}

// Optional phase clocks (-DIPC_PHASE_CLOCKS, profiling builds only): thread 0 of every CTA adds the cycles since its previous
// mark to a global counter per phase. The default build compiles every mark to nothing.
#ifdef IPC_PHASE_CLOCKS
static __device__ unsigned long long g_phase[16];
#define IPC_PH_ARG , long long& ph_last
#define IPC_PH_PASS , ph_last
#ifdef __CUDA_ARCH__
#define IPC_PH(i) do { long long now_ = clock64(); if (threadIdx.x == 0) atomicAdd(&g_phase[i], (unsigned long long)(now_ - ph_last)); ph_last = now_; } while (0)
#define IPC_PH_COUNT(i, n) do { if (threadIdx.x == 0) atomicAdd(&g_phase[i], (unsigned long long)(n)); } while (0)
#else
#define IPC_PH(i) do { (void)ph_last; } while (0)
#define IPC_PH_COUNT(i, n) do {} while (0)
#endif
#else
#define IPC_PH_ARG
#define IPC_PH_PASS
#define IPC_PH(i) do {} while (0)
#define IPC_PH_COUNT(i, n) do {} while (0)
#endif
struct P2 { double x, y, t; };
struct Lin2 {            // linearisation of one relative-pose edge a -> b in its own frame
    double c, s;         // cos / sin of theta_a
    double rx, ry;       // R_a^T (t_b - t_a)
    double d0, d1, d2;   // residual in the relative frame (r - z)
    double w0, w1, w2;   // D d
    double chi;
};
IPC_HD void lin2cs(double c, double s, const P2& a, const P2& b, double zx, double zy, double zt, const double* D, Lin2& e) {
    e.c = c; e.s = s;
    double dx = b.x - a.x, dy = b.y - a.y;
    e.rx = c * dx + s * dy;
    e.ry = -s * dx + c * dy;
    e.d0 = e.rx - zx; e.d1 = e.ry - zy; e.d2 = wrap_pi_hd(b.t - a.t - zt);
    e.w0 = D[0] * e.d0 + D[1] * e.d1 + D[2] * e.d2;
    e.w1 = D[1] * e.d0 + D[3] * e.d1 + D[4] * e.d2;
    e.w2 = D[2] * e.d0 + D[4] * e.d1 + D[5] * e.d2;
    e.chi = e.d0 * e.w0 + e.d1 * e.w1 + e.d2 * e.w2;
}
IPC_HD double quad3(const double* D, double a, double b, double c) {
    return a * (D[0] * a + D[1] * b + D[2] * c) + b * (D[1] * a + D[3] * b + D[4] * c) + c * (D[2] * a + D[4] * b + D[5] * c);
}
// linearised change of the edge residual under vertex increments ha (at a) and hb (at b)
IPC_HD void dlin2(const Lin2& e, const double* ha, const double* hb, double& q0, double& q1, double& q2) {
    double ux = hb[0] - ha[0], uy = hb[1] - ha[1];
    q0 = e.c * ux + e.s * uy + e.ry * ha[2];
    q1 = -e.s * ux + e.c * uy - e.rx * ha[2];
    q2 = hb[2] - ha[2];
}
// gradient pieces of chi2/2 w.r.t. the vertex increments: gi (at a), gj (at b)
IPC_HD void grad2(const Lin2& e, double* gi, double* gj) {
    double rwx = e.c * e.w0 - e.s * e.w1, rwy = e.s * e.w0 + e.c * e.w1;   // R_a w_t
    gj[0] = rwx; gj[1] = rwy; gj[2] = e.w2;
    gi[0] = -rwx; gi[1] = -rwy; gi[2] = e.ry * e.w0 - e.rx * e.w1 - e.w2;
}
IPC_HD void inv_sym3(const double* D, double* V) {
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
IPC_HD void sym3_mul(const double* S, const double* v, double* o) {
    o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}
// C (full 3x3, row-major) = A (sym) * B (sym)
IPC_HD void sym3_sym3(const double* A, const double* B, double* C) {
    const double Af[9] = {A[0], A[1], A[2], A[1], A[3], A[4], A[2], A[4], A[5]};
    const double Bf[9] = {B[0], B[1], B[2], B[1], B[3], B[4], B[2], B[4], B[5]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = Af[r * 3] * Bf[c] + Af[r * 3 + 1] * Bf[3 + c] + Af[r * 3 + 2] * Bf[6 + c];
}

// ------------------------------------------------------------------------------------------------
// block collectives. Host build (tests/host_emul): NT == 1, everything is the identity.
// ------------------------------------------------------------------------------------------------
constexpr int NPRE = 9;      // prefix quantities: PM (00 01 02 11 12 22), Pm (0 1 2)
constexpr int NSPEC = 4;     // special vertices of a check: 0, rs, re, L (all loop end points are among them)
constexpr int SPECW = NPRE + 3;   // published per special vertex: full prefix + pose

// Host build (tests/host_emul, TEST INFRASTRUCTURE): NT = 1 runs a check on the calling thread; NT > 1 emulates the CTA with NT
// OS threads that share a barrier, so the block decomposition (segments, scratch slots, boundary vertices, collectives) is
// exercised on a box without a GPU. The host collectives sum in thread order (the device sums in shuffle-tree order).
struct HostCta { int tid; void (*sync)(void*); void* ctx; };
inline HostCta*& host_cta() { static thread_local HostCta* p = nullptr; return p; }
template <int NT> IPC_HD void bsync() {
#ifdef __CUDA_ARCH__
    if (NT <= 32) __syncwarp(); else __syncthreads();
#else
    if (NT > 1) { HostCta* c = host_cta(); c->sync(c->ctx); }
#endif
}
IPC_HD int hd_tid() {
#ifdef __CUDA_ARCH__
    return threadIdx.x;
#else
    HostCta* c = host_cta();
    return c ? c->tid : 0;
#endif
}

// Exclusive scan over threads of NPRE values + block sums of NS scalars + block max of one value with ONE barrier.
// red: [NW][NPRE + NS + 1] staging (caller double-buffers it).
template <int NT, int NS> struct ScanSumMax {
    static constexpr int W = NPRE + NS + 1;
    static constexpr int NW = NT / 32 > 0 ? NT / 32 : 1;
    IPC_HD static void run(double* v /* in: thread total, out: exclusive prefix */, double* s, double& mx, double* red) {
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        double inc[NPRE];
        if (NT <= 32) {
        // one-warp kernels: one code copy per shuffle distance (the distance loop is NOT unrolled): a fifth of the instructions to fetch —
        // this code runs once per sweep, straight through, bound by instruction fetch (ncu: no_instruction is the top stall of the one-warp
        // kernels) — and the independent values of one distance still pipeline. Measured +1.6 % (profiles/r02_ab_log.txt)
#pragma unroll
        for (int m = 0; m < NPRE; ++m) inc[m] = v[m];
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int m = 0; m < NPRE; ++m) { const double y = __shfl_up_sync(0xffffffffu, inc[m], o); if (lane >= o) inc[m] += y; }
#pragma unroll
            for (int m = 0; m < NS; ++m) s[m] += __shfl_xor_sync(0xffffffffu, s[m], o);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        } else {
#pragma unroll
        for (int m = 0; m < NPRE; ++m) {
            double x = v[m];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { double y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            inc[m] = x;
        }
#pragma unroll
        for (int m = 0; m < NS; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[m] += __shfl_xor_sync(0xffffffffu, s[m], o);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (NT <= 32) {
#pragma unroll
            for (int m = 0; m < NPRE; ++m) v[m] = inc[m] - v[m];
            return;
        }
        if (lane == 31) {
#pragma unroll
            for (int m = 0; m < NPRE; ++m) red[w * W + m] = inc[m];
#pragma unroll
            for (int m = 0; m < NS; ++m) red[w * W + NPRE + m] = s[m];
            red[w * W + NPRE + NS] = mx;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < NPRE; ++m) {
            double base = 0;
            for (int i = 0; i < w; ++i) base += red[i * W + m];
            v[m] = base + inc[m] - v[m];
        }
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            double t = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) t += red[i * W + NPRE + m];
            s[m] = t;
        }
        double t = red[NPRE + NS];
#pragma unroll
        for (int i = 1; i < NW; ++i) t = fmax(t, red[i * W + NPRE + NS]);
        mx = t;
#else
        if (NT == 1) { for (int m = 0; m < NPRE; ++m) v[m] = 0; return; }
        static_assert(NT * W <= 16 * (NPRE + 4), "host emulation: staging holds 16 threads");
        const int t = hd_tid();
        for (int m = 0; m < NPRE; ++m) red[t * W + m] = v[m];
        for (int m = 0; m < NS; ++m) red[t * W + NPRE + m] = s[m];
        red[t * W + NPRE + NS] = mx;
        bsync<NT>();
        for (int m = 0; m < NPRE; ++m) { double b = 0; for (int i = 0; i < t; ++i) b += red[i * W + m]; v[m] = b; }
        for (int m = 0; m < NS; ++m) { double b = 0; for (int i = 0; i < NT; ++i) b += red[i * W + NPRE + m]; s[m] = b; }
        double b = red[NPRE + NS];
        for (int i = 1; i < NT; ++i) b = fmax(b, red[i * W + NPRE + NS]);
        mx = b;
#endif
    }
};

// plain block sum of M values (slow-path sweeps): two barriers, result in every thread
template <int NT, int M> IPC_HD void hd_block_sum(double* v, double* red) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int m = 0; m < M; ++m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[m] += __shfl_xor_sync(0xffffffffu, v[m], o);
    }
    if (NT <= 32) return;
    const int w = threadIdx.x >> 5, NW = NT / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) red[w * M + m] = v[m];
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double s = 0;
        for (int i = 0; i < NW; ++i) s += red[i * M + m];
        v[m] = s;
    }
#else
    if (NT == 1) return;
    const int t = hd_tid();
    bsync<NT>();
    for (int m = 0; m < M; ++m) red[t * M + m] = v[m];
    bsync<NT>();
    for (int m = 0; m < M; ++m) { double b = 0; for (int i = 0; i < NT; ++i) b += red[i * M + m]; v[m] = b; }
#endif
}
// exclusive prefix over threads of M values (dead-reckoning): two barriers
template <int NT, int M> IPC_HD void hd_block_excl_scan(double* v, double* red) {
#ifdef __CUDA_ARCH__
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, NW = NT / 32;
    double inc[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double x = v[m];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        inc[m] = x;
    }
    if (NT > 32) {
        __syncthreads();
        if (lane == 31) {
#pragma unroll
            for (int m = 0; m < M; ++m) red[w * M + m] = inc[m];
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double base = 0;
        if (NT > 32) { for (int i = 0; i < NW; ++i) if (i < w) base += red[i * M + m]; }
        v[m] = base + inc[m] - v[m];
    }
#else
    if (NT == 1) { for (int m = 0; m < M; ++m) v[m] = 0; return; }
    const int t = hd_tid();
    bsync<NT>();
    for (int m = 0; m < M; ++m) red[t * M + m] = v[m];
    bsync<NT>();
    for (int m = 0; m < M; ++m) { double b = 0; for (int i = 0; i < t; ++i) b += red[i * M + m]; v[m] = b; }
#endif
}

// ------------------------------------------------------------------------------------------------
// per-check working set
// ------------------------------------------------------------------------------------------------
enum { STEP_NONE = 0, STEP_GN = 1, STEP_BLEND = 2 };

struct StepSpec {            // GN solution of one linearisation (uniform; lives in shared memory)
    double z[3][3];          // force per region
    double C[3][3];          // constant per region
    double model;            // sum of linearised chi2 after the GN step (diagnostic)
    double gain_loops;       // loop-edge part of h_gn^T H h_gn = sum over edges |J h_gn|^2_Omega (the GN step's predicted gain)
    int rs, re;              // region boundaries (local vertex / edge indices)
};


struct CheckGeom {           // geometry of one check (uniform; lives in the UniBlock)
    int K, lo, L;
    int rs, re, first_is_c, last_is_c;
    int spec_v[NSPEC];       // 0, rs, re, L
    // end points of the loop intervals: every start is 0 or rs, every end is re or L
    int c_a_is_rs, c_b_is_L, m_a_is_rs, m_b_is_L;
};

struct UniBlock {            // uniform per-check data (shared memory): read by every thread, written by thread 0
    LoopRec2 Lc, Lm;
    StepSpec sol;
    double n_c, n_m;         // chi2 of the loop edges at the state of the last sweep
    double lt[2][20];        // per loop (candidate, member): terms t(9), sigma, chi, d(3), cos, sin, x_to, y_to — staged by warp 0
    double sc[12];           // thread-uniform Dogleg scalars that are rarely touched (every thread writes the same values): keeps
                             // them out of the registers that are live across the sweep
    CheckGeom g;
};
constexpr int UNI_DOUBLES = (sizeof(UniBlock) + 7) / 8;

constexpr int RED_DOUBLES_ = 16 * (NPRE + 4);
constexpr int RED2_DOUBLES_ = 16 * 4;            // 16 warps x up to 4 values
struct ChainMem {            // three base pointers + a capacity: cheap to keep in registers and to pass by value
    double* st;              // per-vertex state, AoS of 5 doubles: x, y, theta, cos, sin. Shared memory (MODE 0) or global scratch (MODE 1)
    double* stw;             // where a trial sweep writes the new state. Shared-memory state: == st (in place, the old poses go to the backup B).
                             // Global-memory state: a second buffer — an accepted trial swaps st / stw, a rejected one costs nothing (no backup
                             // stream, no rollback pass)
    double* scr;             // per-CTA global scratch (L2 resident): pose backup AoS[3] x capg, (b, h_gn) AoS[6] x capg, odometry records, indexed
                             // by SLOT (vslot below), not by vertex: the threads of a warp walk their segments in lock step, so
                             // slot = step * NT + thread makes every scratch access of a warp one contiguous run of records
    double* small;           // shared memory: collective staging (2 buffers), special-vertex table (2 buffers), UniBlock
    double* ring;            // shared memory, global-state kernels: per-thread prefetch ring (TileFeed), RING_D x RING_W x NT doubles
    int capv, capg;          // capacity of the state arrays (vertices) and of the scratch arrays (slots)
    IPC_HD double* P(int j) const { return st + 5 * j; }                          // x y theta cos sin of vertex j (state in shared memory)
    IPC_HD double* B(int sl) const { return scr + 3 * sl; }                       // pose backup: state before the last trial sweep
    IPC_HD double* G(int sl) const { return scr + 3 * (size_t)capg + 6 * sl; }    // gradient b_j (3) and h_gn,j (3), g2o vertex coordinates
    IPC_HD double* Z() const { return scr + 9 * (size_t)capg; }                  // odometry records of the window, slot order (3 or 9 / edge)
    IPC_HD double* red() const { return small; }                                 // sweep collective (ScanSumMax), double-buffered by the caller
    IPC_HD double* red2() const { return small + 2 * RED_DOUBLES_; }             // two-barrier collectives (block sum, dead-reckoning scans): their
                                                                                 // own staging, so a sweep that follows without a barrier cannot
                                                                                 // overwrite values a slower warp is still reading
    IPC_HD double* spec() const { return small + 2 * RED_DOUBLES_ + RED2_DOUBLES_; }
    IPC_HD UniBlock* U() const { return reinterpret_cast<UniBlock*>(small + 2 * RED_DOUBLES_ + RED2_DOUBLES_ + 2 * NSPEC * SPECW); }
};
// State record of a vertex (x y theta cos sin). Shared-memory state (GST = false): AoS by VERTEX, components 1 double apart (odd segment
// lengths keep the strided accesses of a warp conflict free). Global-memory state (GST = true: windows that do not fit shared memory,
// and the one-warp-per-check kernels): tiles of 5 x NT doubles by STEP — the vertex thread t reaches at step i of its segment walk
// (vertex k0(t) + 1 + i) is element t of tile i + 1, components NT doubles apart — so every access of a warp is one contiguous
// 256-byte run (coalesced in L2 / HBM). Tile 0 holds the fixed origin (vertex 0) in element 0.
template <int NT, bool GST> struct StateAt {
    static constexpr int CS = GST ? NT : 1;                                       // distance between the components of one record
    IPC_HD static double* step(const ChainMem& M, int j, int i, int t) {         // vertex j = k0(t) + 1 + i of thread t
        return GST ? M.st + ((size_t)(i + 1) * 5) * NT + t : M.st + 5 * j;
    }
    IPC_HD static double* stepw(const ChainMem& M, int j, int i, int t) {        // same record in the buffer a trial sweep writes
        return GST ? M.stw + ((size_t)(i + 1) * 5) * NT + t : M.stw + 5 * j;
    }
    IPC_HD static double* vertex(const ChainMem& M, int j, int S) {              // any vertex (rare accesses: loop end points)
        if (!GST) return M.st + 5 * j;
        if (j == 0) return M.st;
        const int t = (j - 1) / S, i = (j - 1) - t * S;
        return M.st + ((size_t)(i + 1) * 5) * NT + t;
    }
};
IPC_HD constexpr int global_state_doubles(int capv, int nt) { return 5 * (capv + 4 * nt); }   // tiles: (S + 1) * 5 * NT <= 5 (L + 3 NT)
// scratch slot of vertex j for segments of S vertices per thread: vertex 0 -> 0; vertex k0 + 1 + i of thread t -> i * NT + t + 1.
// Slots reach S * NT <= L + 2 NT, hence capg = capv + 2 NT + 2.
template <int NT> IPC_HD int vslot(int j, int S) { return j == 0 ? 0 : ((j - 1) % S) * NT + (j - 1) / S + 1; }
template <int NT> IPC_HD constexpr int scratch_slots(int capv) { return capv + 2 * NT + 2; }
constexpr int RED_DOUBLES = 16 * (NPRE + 4);     // NW <= 16 warps (NT <= 512) x (NPRE + NS + 1), NS = 3
constexpr int CHAIN_SMALL_DOUBLES = 2 * RED_DOUBLES + RED2_DOUBLES_ + 2 * NSPEC * SPECW + UNI_DOUBLES + (UNI_DOUBLES & 1);
IPC_HD constexpr int stage_doubles(int capv) { return (3 * (capv + 2) + 1) & ~1; }   // MODE 2: staged odometry records (UNI, 3 doubles / edge), even
constexpr int CHAIN_STATE_ARRAYS = 5;            // per-vertex doubles in shared memory (MODE 0)
constexpr int CHAIN_SCRATCH_ARRAYS = 18;         // per-slot doubles in the global scratch (backup 3, b + h_gn 6, odometry record <= 9)


struct OdomView {            // odometry records of the window in HBM/L2, AoS: (zx zy zt) when every edge shares one isotropic
    const double* rec;       // information (UNI, 24 B / edge), else (zx zy zt d00 d01 d02 d11 d12 d22) (72 B / edge)
    const double* Du;        // UNI: the shared information matrix (frame independent) ...
    const double* Vu;        //      ... and its inverse
    double* zs;              // the window's records again, in the per-CTA scratch in SLOT order (edge k0 + i of thread t at i * NT + t):
                             // written once per check by the dead-reckoning pass, read by every later pass as contiguous warp accesses
    const double* sm;        // STG kernels: the window's records in SHARED memory, natural order (record of local edge k at sm + 3 k),
                             // landed there by ONE bulk asynchronous copy (cp.async.bulk + mbarrier) per check
};
template <bool UNI> IPC_HD const double* odom_rec(const OdomView& O, int k) { return O.rec + (UNI ? 3 : 9) * (size_t)k; }
IPC_HD double ldg_d(const double* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

struct ThreadState {         // registers carried from sweep to sweep
    int k0, k1;              // owned edges [k0, k1); owned vertices k0+1 .. k1
    int S, tid;              // segment length (vertices per thread) and the thread's index: scratch slots (vslot)
    P2 pa; double ca, sa;    // pose (+ cos / sin) of vertex k0 at the current state
    double base[NPRE];       // prefix (PM, Pm) at vertex k0 for the linearisation of the current state
};

// twist Xi_j = Pm_j - PM_j z_r - C_r of the GN step at vertex j (region r by position)
IPC_HD void twist_at(const StepSpec* sp, int j, const double* pre, double* Xi) {
    const int r = (j <= sp->rs) ? 0 : (j <= sp->re ? 1 : 2);
    const double* z = sp->z[r]; const double* C = sp->C[r];
    const double z0 = z[0], z1 = z[1], z2 = z[2];
    Xi[0] = pre[6] - (pre[0] * z0 + pre[1] * z1 + pre[2] * z2) - C[0];
    Xi[1] = pre[7] - (pre[1] * z0 + pre[3] * z1 + pre[4] * z2) - C[1];
    Xi[2] = pre[8] - (pre[2] * z0 + pre[4] * z1 + pre[5] * z2) - C[2];
}
IPC_HD void gn_step_at(const StepSpec* sp, int j, const double* pre /* full prefix at j */, double x, double y, double* h) {
    double X[3]; twist_at(sp, j, pre, X);
    h[0] = X[0] - y * X[2];      // u = T Xi, T = [I, S t; 0 1]
    h[1] = X[1] + x * X[2];
    h[2] = X[2];
}

// contribution of one linearised edge to the prefix sums: M = Q V Q^T (6), m = -Q d (3); Q = [c -s yb; s c -xb; 0 0 1]
IPC_HD void edge_prefix_terms(const Lin2& e, const double* V, double xb, double yb, double* t) {
    const double q02 = yb, q12 = -xb;
    const double r00 = e.c * V[0] - e.s * V[1] + q02 * V[2], r01 = e.c * V[1] - e.s * V[3] + q02 * V[4], r02 = e.c * V[2] - e.s * V[4] + q02 * V[5];
    const double r10 = e.s * V[0] + e.c * V[1] + q12 * V[2], r11 = e.s * V[1] + e.c * V[3] + q12 * V[4], r12 = e.s * V[2] + e.c * V[4] + q12 * V[5];
    t[0] = r00 * e.c - r01 * e.s + r02 * q02;
    t[1] = r00 * e.s + r01 * e.c + r02 * q12;
    t[2] = r02;
    t[3] = r10 * e.s + r11 * e.c + r12 * q12;
    t[4] = r12;
    t[5] = V[5];
    t[6] = -(e.c * e.d0 - e.s * e.d1 + q02 * e.d2);
    t[7] = -(e.s * e.d0 + e.c * e.d1 + q12 * e.d2);
    t[8] = -e.d2;
}
// same for V = diag(va, va, vc) (uniform isotropic information): Q V Q^T no longer depends on the heading
IPC_HD void edge_prefix_terms_iso(const Lin2& e, double va, double vc, double xb, double yb, double* t) {
    const double vy = vc * yb, vx = vc * xb;
    t[0] = fma(vy, yb, va);
    t[1] = -vy * xb;
    t[2] = vy;
    t[3] = fma(vx, xb, va);
    t[4] = -vx;
    t[5] = vc;
    t[6] = -(e.c * e.d0 - e.s * e.d1 + yb * e.d2);
    t[7] = -(e.s * e.d0 + e.c * e.d1 - xb * e.d2);
    t[8] = -e.d2;
}
// z[0..2] (and D for the general case) of one odometry edge, loaded ahead of use
template <bool UNI> struct OdomRec { double z[UNI ? 3 : 9]; };
template <bool UNI, bool STG = false> IPC_HD void odom_load(const OdomView& O, int es /* edge slot */, int k /* local edge */, OdomRec<UNI>& r) {
    const double* p = STG ? O.sm + 3 * (size_t)k : O.zs + (UNI ? 3 : 9) * (size_t)es;
#pragma unroll
    for (int q = 0; q < (UNI ? 3 : 9); ++q) r.z[q] = p[q];     // plain loads: this kernel wrote them
}
template <bool UNI> IPC_HD void odom_terms(const OdomView& O, const OdomRec<UNI>& r, double c, double s, const P2& a, const P2& b, Lin2& e, double* t) {
    if (UNI) {
        lin2cs(c, s, a, b, r.z[0], r.z[1], r.z[2], O.Du, e);
        edge_prefix_terms_iso(e, O.Vu[0], O.Vu[5], b.x, b.y, t);
    } else {
        double V[6];
        inv_sym3(r.z + 3, V);
        lin2cs(c, s, a, b, r.z[0], r.z[1], r.z[2], r.z + 3, e);
        edge_prefix_terms(e, V, b.x, b.y, t);
    }
}

struct SweepOut { double chi, mx, hh, gain; };   // odometry chi2 sum / max at the new state, |h|^2 of the applied step, odometry part of h^T H h (GN)

// The sweep: apply a step (none / GN of M.U()->sol / blend c1 b + c2 h_gn from the scratch), re-linearise, chi2, interval
// sums of the new linearisation at the special vertices. The GN step at vertex j needs the prefix of the OLD linearisation at
// j: it is rebuilt on the fly from the old poses (no sincos: cos / sin are stored). Writes the pose backup when a step is
// applied. Two block barriers.
//
// One edge of the sweep is two dependent halves: A(k) = old linearisation of edge k + step of vertex k + 1 (-> its new pose),
// B(k) = sincos of the new heading + new linearisation of edge k + the accumulations. B(k) and A(k + 1) do not depend on each
// other (A only reads OLD poses and the running old prefix), so the loop is software pipelined by hand: iteration k holds
// B(k) and A(k + 1) in ONE basic block (one loop per step kind, no branches in the body, loads first, stores last), which lets the
// scheduler interleave the two dependent chains — the edge loop is bound by dependent-issue latency, not by the fp64 pipe.
// Every accumulation keeps its order (k ascending), so the results are bit-identical to the unpipelined loop.
template <bool UNI> IPC_HD void sweep_gn_step(const OdomView& O, const OdomRec<UNI>& r, const double* zr, const double* Cr, const P2& oa, double oca, double osa,
                                              const P2& ob, double* pre, double& gain, double* h) {
    Lin2 eo; double to[NPRE];
    odom_terms<UNI>(O, r, oca, osa, oa, ob, eo, to);
    const double z0 = zr[0], z1 = zr[1], z2 = zr[2];
    {   // predicted gain h_gn^T H h_gn, accumulated edge by edge from non-negative terms |J h_gn|^2_Omega (chi2 - model
        // cancels catastrophically near convergence and for gross outliers): the residual change of edge k under the
        // force z of its region is -(d + V Q^T z)
        const double y0 = eo.c * z0 + eo.s * z1, y1 = -eo.s * z0 + eo.c * z1, y2 = ob.y * z0 - ob.x * z1 + z2;
        if (UNI) {
            const double w0 = fma(O.Vu[0], y0, eo.d0), w1 = fma(O.Vu[0], y1, eo.d1), w2 = fma(O.Vu[5], y2, eo.d2);
            gain += O.Du[0] * (w0 * w0 + w1 * w1) + O.Du[5] * w2 * w2;
        } else {
            double V6[6]; inv_sym3(r.z + (UNI ? 0 : 3), V6);
            const double w0 = eo.d0 + V6[0] * y0 + V6[1] * y1 + V6[2] * y2;
            const double w1 = eo.d1 + V6[1] * y0 + V6[3] * y1 + V6[4] * y2;
            const double w2 = eo.d2 + V6[2] * y0 + V6[4] * y1 + V6[5] * y2;
            gain += quad3(r.z + (UNI ? 0 : 3), w0, w1, w2);
        }
    }
#pragma unroll
    for (int m = 0; m < NPRE; ++m) pre[m] += to[m];
    // twist Xi = Pm - PM z - C of the vertex's region, u = T Xi (gn_step_at with the region's z, C already in registers)
    const double X0 = pre[6] - (pre[0] * z0 + pre[1] * z1 + pre[2] * z2) - Cr[0];
    const double X1 = pre[7] - (pre[1] * z0 + pre[3] * z1 + pre[4] * z2) - Cr[1];
    const double X2 = pre[8] - (pre[2] * z0 + pre[4] * z1 + pre[5] * z2) - Cr[2];
    h[0] = X0 - ob.y * X2; h[1] = X1 + ob.x * X2; h[2] = X2;
}

// Shared-memory state (the CTA-per-check kernels): the state is read in place, odometry records and gradients one step ahead
// through registers.
template <int NT, bool UNI, bool STG, int STEP, bool GST = false> IPC_HD void sweep_mode_direct(const ChainMem& M, const OdomView& O, double c1, double c2, ThreadState& ts,
                                                                       SweepOut& out, int& buf, const int* spec_v IPC_PH_ARG) {
    using SA = StateAt<NT, GST>;
    constexpr int CS = SA::CS;
    const int k0 = ts.k0, k1 = ts.k1;
    const StepSpec* sp = &M.U()->sol;
    double* spec = M.spec() + (size_t)buf * NSPEC * SPECW;
    double pre[NPRE];        // running prefix of the OLD linearisation (GN mode)
#pragma unroll
    for (int m = 0; m < NPRE; ++m) pre[m] = ts.base[m];
    P2 oa = ts.pa; double oca = ts.ca, osa = ts.sa;      // old from-vertex of the next A
    P2 na = oa; double nca = oca, nsa = osa;             // new from-vertex of the next B
    if (STEP != STEP_NONE && k0 > 0 && k0 < k1) {        // boundary vertex k0: same arithmetic as its owner => identical bits
        double h[3];
        if (STEP == STEP_GN) gn_step_at(sp, k0, pre, oa.x, oa.y, h);
        else {
            const double* gq = M.G((ts.S - 1) * NT + ts.tid);      // vertex k0 = the last vertex of thread tid - 1
#pragma unroll
            for (int q = 0; q < 3; ++q) h[q] = c1 * gq[q] + c2 * gq[3 + q];
        }
        na.x += h[0]; na.y += h[1]; na.t = wrap_pi_hd(na.t + h[2]);
        ipc_sincos(na.t, &nsa, &nca);
    }
    ts.pa = na; ts.ca = nca; ts.sa = nsa;
    double run[NPRE];        // local prefix of the NEW linearisation
#pragma unroll
    for (int m = 0; m < NPRE; ++m) run[m] = 0;
    double chi = 0, mx = 0, hh = 0, gain = 0;
    bool has_spec = false;
#pragma unroll
    for (int q = 1; q < NSPEC; ++q) has_spec |= (spec_v[q] > k0 && spec_v[q] <= k1);
    const int v_rs = spec_v[1], v_re = spec_v[2];        // region boundaries (sp->rs, sp->re)
    if (k0 < k1) {
        int sl = ts.tid + 1;                     // scratch slot of vertex k + 1 (the vertex B(k) finishes); + NT per vertex
        // odometry records: rB of edge k (for B), rA of edge k + 1 (for A), loaded one more iteration ahead (L2 latency)
        OdomRec<UNI> rB, rA;
        double gA[6] = {0, 0, 0, 0, 0, 0};       // blend steps: (b, h_gn) of the vertex the next A moves, from the global scratch
        odom_load<UNI, STG>(O, sl - 1, k0, rB);                    // edge slot = slot of the edge's head vertex - 1
        {
            const bool more = k0 + 1 < k1;
            odom_load<UNI, STG>(O, more ? sl + NT - 1 : sl - 1, more ? k0 + 1 : k0, rA);
            if (STEP == STEP_BLEND) { const double* gq = M.G(more ? sl + NT : sl);
#pragma unroll
                for (int q = 0; q < 6; ++q) gA[q] = gq[q]; }
        }
        // ---- prologue: A(k0) ----
        P2 nb; double ncb, nsb;                  // vertex k + 1 at the new state (cos / sin: known only for STEP_NONE before B)
        {
            const double* pq0 = SA::step(M, k0 + 1, 0, ts.tid);
            const P2 ob{pq0[0], pq0[CS], pq0[2 * CS]};
            const double ocb = pq0[3 * CS], osb = pq0[4 * CS];
            nb = ob; ncb = ocb; nsb = osb;
            if (STEP != STEP_NONE) {
                double h[3];
                if (STEP == STEP_GN) {
                    const int rg = (k0 < v_rs) ? 0 : (k0 < v_re ? 1 : 2);
                    sweep_gn_step<UNI>(O, rB, sp->z[rg], sp->C[rg], oa, oca, osa, ob, pre, gain, h);
                } else {
                    const double* gq = M.G(sl);
#pragma unroll
                    for (int q = 0; q < 3; ++q) h[q] = c1 * gq[q] + c2 * gq[3 + q];
                }
                double* bq = M.B(sl);
                bq[0] = ob.x; bq[1] = ob.y; bq[2] = ob.t;
                nb.x += h[0]; nb.y += h[1]; nb.t = wrap_pi_hd(nb.t + h[2]);
                hh += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
            }
            oa = ob; oca = ocb; osa = osb;
        }
        // ---- steady state: B(k) next to A(k + 1) ----
        int k = k0;
        // GST (state in the global scratch): the old pose of vertex k + 2 is fetched one iteration ahead as well (L2 / HBM latency)
        double pn[5] = {0, 0, 0, 0, 0};
        if (GST) { const int jn = k0 + 2 <= k1 ? k0 + 2 : k1; const double* pq = SA::step(M, jn, jn - k0 - 1, ts.tid);
#pragma unroll
            for (int q = 0; q < 5; ++q) pn[q] = pq[q * CS]; }
        for (; k + 1 < k1; ++k, sl += NT) {
            // loads first (shared memory: old pose of vertex k + 2, the region's force; global: records / gradients two edges ahead)
            const int jn = GST ? (k + 3 <= k1 ? k + 3 : k1) : k + 2;
            const double* pq2 = SA::step(M, jn, jn - k0 - 1, ts.tid);
            double pc5[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) { pc5[q] = GST ? pn[q] : pq2[q * CS]; if (GST) pn[q] = pq2[q * CS]; }
            const P2 ob2{pc5[0], pc5[1], pc5[2]};
            const double ocb2 = pc5[3], osb2 = pc5[4];
            double zr[3] = {0, 0, 0}, Cr[3] = {0, 0, 0};
            if (STEP == STEP_GN) {
                const int rg = (k + 1 < v_rs) ? 0 : (k + 1 < v_re ? 1 : 2);
#pragma unroll
                for (int q = 0; q < 3; ++q) { zr[q] = sp->z[rg][q]; Cr[q] = sp->C[rg][q]; }
            }
            const bool more = k + 2 < k1;
            OdomRec<UNI> rP;
            odom_load<UNI, STG>(O, more ? sl + 2 * NT - 1 : sl + NT - 1, more ? k + 2 : k + 1, rP);
            double gP[6] = {0, 0, 0, 0, 0, 0};
            if (STEP == STEP_BLEND) { const double* gq = M.G(more ? sl + 2 * NT : sl + NT);
#pragma unroll
                for (int q = 0; q < 6; ++q) gP[q] = gq[q]; }
            // B(k): vertex j = k + 1 gets its cos / sin, edge k its new linearisation
            const int j = k + 1;
            if (STEP != STEP_NONE) ipc_sincos(nb.t, &nsb, &ncb);
            Lin2 e; double t[NPRE];
            odom_terms<UNI>(O, rB, nca, nsa, na, nb, e, t);
            chi += e.chi; mx = fmax(mx, e.chi);
#pragma unroll
            for (int m = 0; m < NPRE; ++m) run[m] += t[m];
            // A(k + 1): step of vertex k + 2 from the old linearisation of edge k + 1
            P2 nb2 = ob2;
            if (STEP != STEP_NONE) {
                double h[3];
                if (STEP == STEP_GN) sweep_gn_step<UNI>(O, rA, zr, Cr, oa, oca, osa, ob2, pre, gain, h);
                else {
#pragma unroll
                    for (int q = 0; q < 3; ++q) h[q] = c1 * gA[q] + c2 * gA[3 + q];
                }
                nb2.x += h[0]; nb2.y += h[1]; nb2.t = wrap_pi_hd(nb2.t + h[2]);
                hh += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
            }
            // stores last
            if (STEP != STEP_NONE) {
                double* pq = SA::step(M, j, k - k0, ts.tid);
                pq[0] = nb.x; pq[CS] = nb.y; pq[2 * CS] = nb.t; pq[3 * CS] = ncb; pq[4 * CS] = nsb;
                double* bq = M.B(sl + NT);
                bq[0] = ob2.x; bq[1] = ob2.y; bq[2] = ob2.t;
            }
            if (has_spec) {
#pragma unroll
                for (int q = 1; q < NSPEC; ++q) {
                    if (j == spec_v[q]) {        // local part now, the thread base is added after the scan
                        double* o = spec + q * SPECW;
#pragma unroll
                        for (int m = 0; m < NPRE; ++m) o[m] = run[m];
                        o[NPRE] = nb.x; o[NPRE + 1] = nb.y; o[NPRE + 2] = nb.t;
                    }
                }
            }
            na = nb; nca = ncb; nsa = nsb;
            nb = nb2; ncb = ocb2; nsb = osb2;
            oa = ob2; oca = ocb2; osa = osb2;
            rB = rA; rA = rP;
#pragma unroll
            for (int q = 0; q < 6; ++q) gA[q] = gP[q];
        }
        // ---- epilogue: B(k1 - 1) ----
        {
            const int j = k + 1;
            if (STEP != STEP_NONE) {
                ipc_sincos(nb.t, &nsb, &ncb);
                double* pq = SA::step(M, j, k - k0, ts.tid);
                pq[0] = nb.x; pq[CS] = nb.y; pq[2 * CS] = nb.t; pq[3 * CS] = ncb; pq[4 * CS] = nsb;
            }
            Lin2 e; double t[NPRE];
            odom_terms<UNI>(O, rB, nca, nsa, na, nb, e, t);
            chi += e.chi; mx = fmax(mx, e.chi);
#pragma unroll
            for (int m = 0; m < NPRE; ++m) run[m] += t[m];
            if (has_spec) {
#pragma unroll
                for (int q = 1; q < NSPEC; ++q) {
                    if (j == spec_v[q]) {
                        double* o = spec + q * SPECW;
#pragma unroll
                        for (int m = 0; m < NPRE; ++m) o[m] = run[m];
                        o[NPRE] = nb.x; o[NPRE + 1] = nb.y; o[NPRE + 2] = nb.t;
                    }
                }
            }
        }
    }
    double s[3] = {chi, hh, gain};
    IPC_PH(1);
    ScanSumMax<NT, 3>::run(run, s, mx, M.red() + (size_t)buf * RED_DOUBLES);
#pragma unroll
    for (int m = 0; m < NPRE; ++m) ts.base[m] = run[m];
    out.chi = s[0]; out.hh = s[1]; out.gain = s[2]; out.mx = mx;
    if (has_spec) {
#pragma unroll
        for (int q = 1; q < NSPEC; ++q) {
            const int v = spec_v[q];
            if (v > k0 && v <= k1) {
                double* o = spec + q * SPECW;
#pragma unroll
                for (int m = 0; m < NPRE; ++m) o[m] += run[m];
            }
        }
    }
    bsync<NT>();
    buf ^= 1;
    IPC_PH(2);
}

// ---- the per-step input stream of a pass ---------------------------------------------------------------------------------
// Step i of a thread's segment walk needs the state of vertex k0 + 1 + i, the odometry record of edge k0 + i (its head is that
// vertex) and, in blend sweeps, (b, h_gn) of that vertex. TileFeed hands these out in step order:
//   * state in shared memory (GST = false): the state is read in place, record and gradient one step ahead through registers;
//   * state in global memory, uniform-information kernels on the device (RING): everything for step i + RING_D is already on its
//     way into a per-thread ring in shared memory (cp.async, one commit group per step), so the L2 / HBM latency of the streamed
//     window is covered by RING_D steps of arithmetic without holding registers;
//   * state in global memory otherwise (general information, host emulation): one step ahead through registers.
IPC_HD constexpr int ring_depth(int nt) { return nt > 256 ? 2 : 4; }      // steps in flight (the 512-thread ring must still fit shared memory)
constexpr int RING_W = 14;           // doubles per step and thread: state 5, odometry record 3, (b, h_gn) 6
IPC_HD constexpr int ring_doubles(int nt) { return ring_depth(nt) * RING_W * nt; }
#ifndef IPC_RING_ODOM_SHARED
#define IPC_RING_ODOM_SHARED 1
#endif
#ifdef __CUDA_ARCH__
#define IPC_RING_DEVICE 1
#else
#define IPC_RING_DEVICE 0
#endif
template <int NT, bool UNI, bool STG, bool GST, bool WITH_G> struct TileFeed {
    static constexpr bool RING = GST && UNI && !STG && (IPC_RING_DEVICE != 0);
    static constexpr int RING_D = ring_depth(NT);
    static constexpr int CS = StateAt<NT, GST>::CS;
    const ChainMem& M; const OdomView& O;
    int k0, n, tid;                      // first edge, number of steps, thread
    double bst[5]; OdomRec<UNI> brec; double bg[6];      // register look-ahead (paths without the ring)
    IPC_HD TileFeed(const ChainMem& M_, const OdomView& O_, int k0_, int n_, int tid_) : M(M_), O(O_), k0(k0_), n(n_), tid(tid_) {}
    IPC_HD void load_direct(int i, double* st5, OdomRec<UNI>& rec, double* g6) const {
        const int ic = i < n ? i : n - 1;                // past the end: a valid address, the values are not used
        if (GST) { const double* ps = StateAt<NT, GST>::step(M, k0 + 1 + ic, ic, tid);
#pragma unroll
            for (int q = 0; q < 5; ++q) st5[q] = ps[q * CS]; }
        odom_load<UNI, STG>(O, ic * NT + tid, k0 + ic, rec);
        if (WITH_G) { const double* gq = M.G((ic * NT + tid) + 1);
#pragma unroll
            for (int q = 0; q < 6; ++q) g6[q] = gq[q]; }
    }
#ifdef __CUDA_ARCH__
    IPC_HD void ring_issue(int i) const {
        if (i < n) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(M.ring) + (unsigned)((((i & (RING_D - 1)) * RING_W) * NT + tid) * 8);
            const double* ps = StateAt<NT, GST>::step(M, 0, i, tid);
#pragma unroll
            for (int q = 0; q < 5; ++q) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + q * NT * 8), "l"(ps + q * CS) : "memory");
            // odometry record of edge k0 + i: from the graph's own array (84 KB for M3500: L1 / L2 resident, shared by every check) —
            // 24 scattered bytes per thread, but no per-check copy of the window to stream from HBM
            const double* po = IPC_RING_ODOM_SHARED ? O.rec + 3 * (size_t)(k0 + i) : O.zs + 3 * (size_t)(i * NT + tid);
#pragma unroll
            for (int q = 0; q < 3; ++q) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (5 + q) * NT * 8), "l"(po + q) : "memory");
            if (WITH_G) { const double* gq = M.G((i * NT + tid) + 1);
#pragma unroll
                for (int q = 0; q < 6; ++q) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (8 + q) * NT * 8), "l"(gq + q) : "memory"); }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#endif
    IPC_HD void start() {
#ifdef __CUDA_ARCH__
        if (RING) {
#pragma unroll
            for (int i = 0; i < RING_D; ++i) ring_issue(i);
            return;
        }
#endif
        load_direct(0, bst, brec, bg);
    }
    // inputs of step i (called for i = 0, 1, 2, ... in order; i may run past the end by one: clamped)
    IPC_HD void next(int i, double* st5, OdomRec<UNI>& rec, double* g6) {
#ifdef __CUDA_ARCH__
        if (RING) {
            asm volatile("cp.async.wait_group %0;" ::"n"(RING_D - 1) : "memory");
            const double* src = M.ring + (((i & (RING_D - 1)) * RING_W) * NT + tid);
#pragma unroll
            for (int q = 0; q < 5; ++q) st5[q] = src[q * NT];
#pragma unroll
            for (int q = 0; q < 3; ++q) rec.z[q] = src[(5 + q) * NT];
            if (WITH_G) {
#pragma unroll
                for (int q = 0; q < 6; ++q) g6[q] = src[(8 + q) * NT]; }
            ring_issue(i + RING_D);
            return;
        }
#endif
        if (GST) {
#pragma unroll
            for (int q = 0; q < 5; ++q) st5[q] = bst[q]; }
        else { const int ic = i < n ? i : n - 1; const double* ps = StateAt<NT, GST>::step(M, k0 + 1 + ic, ic, tid);
#pragma unroll
            for (int q = 0; q < 5; ++q) st5[q] = ps[q * CS]; }
        rec = brec;
        if (WITH_G) {
#pragma unroll
            for (int q = 0; q < 6; ++q) g6[q] = bg[q]; }
        load_direct(i + 1, bst, brec, bg);
    }
    IPC_HD void finish() const {
#ifdef __CUDA_ARCH__
        if (RING) asm volatile("cp.async.wait_group 0;" ::: "memory");     // the ring is reused by the next pass
#endif
    }
};

template <int NT, bool UNI, bool STG, int STEP, bool GST = false> IPC_HD void sweep_mode(const ChainMem& M, const OdomView& O, double c1, double c2, ThreadState& ts,
                                                                       SweepOut& out, int& buf, const int* spec_v IPC_PH_ARG) {
    using SA = StateAt<NT, GST>;
    constexpr int CS = SA::CS;
    const int k0 = ts.k0, k1 = ts.k1;
    const StepSpec* sp = &M.U()->sol;
    double* spec = M.spec() + (size_t)buf * NSPEC * SPECW;
    double pre[NPRE];        // running prefix of the OLD linearisation (GN mode)
#pragma unroll
    for (int m = 0; m < NPRE; ++m) pre[m] = ts.base[m];
    P2 oa = ts.pa; double oca = ts.ca, osa = ts.sa;      // old from-vertex of the next A
    P2 na = oa; double nca = oca, nsa = osa;             // new from-vertex of the next B
    if (STEP != STEP_NONE && k0 > 0 && k0 < k1) {        // boundary vertex k0: same arithmetic as its owner => identical bits
        double h[3];
        if (STEP == STEP_GN) gn_step_at(sp, k0, pre, oa.x, oa.y, h);
        else {
            const double* gq = M.G((ts.S - 1) * NT + ts.tid);      // vertex k0 = the last vertex of thread tid - 1
#pragma unroll
            for (int q = 0; q < 3; ++q) h[q] = c1 * gq[q] + c2 * gq[3 + q];
        }
        na.x += h[0]; na.y += h[1]; na.t = wrap_pi_hd(na.t + h[2]);
        ipc_sincos(na.t, &nsa, &nca);
    }
    ts.pa = na; ts.ca = nca; ts.sa = nsa;
    double run[NPRE];        // local prefix of the NEW linearisation
#pragma unroll
    for (int m = 0; m < NPRE; ++m) run[m] = 0;
    double chi = 0, mx = 0, hh = 0, gain = 0;
    bool has_spec = false;
#pragma unroll
    for (int q = 1; q < NSPEC; ++q) has_spec |= (spec_v[q] > k0 && spec_v[q] <= k1);
    const int v_rs = spec_v[1], v_re = spec_v[2];        // region boundaries (sp->rs, sp->re)
    if (k0 < k1) {
        int sl = ts.tid + 1;                     // scratch slot of vertex k + 1 (the vertex B(k) finishes); + NT per vertex
        TileFeed<NT, UNI, STG, GST, STEP == STEP_BLEND> feed(M, O, k0, k1 - k0, ts.tid);
        feed.start();
        OdomRec<UNI> rB;                         // record of edge k (B uses it one step after A did)
        // ---- prologue: A(k0) ----
        P2 nb; double ncb, nsb;                  // vertex k + 1 at the new state (cos / sin: known only for STEP_NONE before B)
        {
            double p5[5], g6[6];
            feed.next(0, p5, rB, g6);
            const P2 ob{p5[0], p5[1], p5[2]};
            const double ocb = p5[3], osb = p5[4];
            nb = ob; ncb = ocb; nsb = osb;
            if (STEP != STEP_NONE) {
                double h[3];
                if (STEP == STEP_GN) {
                    const int rg = (k0 < v_rs) ? 0 : (k0 < v_re ? 1 : 2);
                    sweep_gn_step<UNI>(O, rB, sp->z[rg], sp->C[rg], oa, oca, osa, ob, pre, gain, h);
                } else {
#pragma unroll
                    for (int q = 0; q < 3; ++q) h[q] = c1 * g6[q] + c2 * g6[3 + q];
                }
                if (!GST) { double* bq = M.B(sl); bq[0] = ob.x; bq[1] = ob.y; bq[2] = ob.t; }
                nb.x += h[0]; nb.y += h[1]; nb.t = wrap_pi_hd(nb.t + h[2]);
                hh += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
            }
            oa = ob; oca = ocb; osa = osb;
        }
        // ---- steady state: B(k) next to A(k + 1) ----
        int k = k0;
        for (; k + 1 < k1; ++k, sl += NT) {
            // loads first (old pose of vertex k + 2, record of edge k + 1, the region's force)
            double p5[5], gA[6];
            OdomRec<UNI> rA;
            feed.next(k - k0 + 1, p5, rA, gA);
            const P2 ob2{p5[0], p5[1], p5[2]};
            const double ocb2 = p5[3], osb2 = p5[4];
            double zr[3] = {0, 0, 0}, Cr[3] = {0, 0, 0};
            if (STEP == STEP_GN) {
                const int rg = (k + 1 < v_rs) ? 0 : (k + 1 < v_re ? 1 : 2);
#pragma unroll
                for (int q = 0; q < 3; ++q) { zr[q] = sp->z[rg][q]; Cr[q] = sp->C[rg][q]; }
            }
            // B(k): vertex j = k + 1 gets its cos / sin, edge k its new linearisation
            const int j = k + 1;
            if (STEP != STEP_NONE) ipc_sincos(nb.t, &nsb, &ncb);
            Lin2 e; double t[NPRE];
            odom_terms<UNI>(O, rB, nca, nsa, na, nb, e, t);
            chi += e.chi; mx = fmax(mx, e.chi);
#pragma unroll
            for (int m = 0; m < NPRE; ++m) run[m] += t[m];
            // A(k + 1): step of vertex k + 2 from the old linearisation of edge k + 1
            P2 nb2 = ob2;
            if (STEP != STEP_NONE) {
                double h[3];
                if (STEP == STEP_GN) sweep_gn_step<UNI>(O, rA, zr, Cr, oa, oca, osa, ob2, pre, gain, h);
                else {
#pragma unroll
                    for (int q = 0; q < 3; ++q) h[q] = c1 * gA[q] + c2 * gA[3 + q];
                }
                nb2.x += h[0]; nb2.y += h[1]; nb2.t = wrap_pi_hd(nb2.t + h[2]);
                hh += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
            }
            // stores last
            if (STEP != STEP_NONE) {
                double* pq = SA::stepw(M, j, k - k0, ts.tid);
                pq[0] = nb.x; pq[CS] = nb.y; pq[2 * CS] = nb.t; pq[3 * CS] = ncb; pq[4 * CS] = nsb;
                if (!GST) { double* bq = M.B(sl + NT); bq[0] = ob2.x; bq[1] = ob2.y; bq[2] = ob2.t; }
            }
            if (has_spec) {
#pragma unroll
                for (int q = 1; q < NSPEC; ++q) {
                    if (j == spec_v[q]) {        // local part now, the thread base is added after the scan
                        double* o = spec + q * SPECW;
#pragma unroll
                        for (int m = 0; m < NPRE; ++m) o[m] = run[m];
                        o[NPRE] = nb.x; o[NPRE + 1] = nb.y; o[NPRE + 2] = nb.t;
                    }
                }
            }
            na = nb; nca = ncb; nsa = nsb;
            nb = nb2; ncb = ocb2; nsb = osb2;
            oa = ob2; oca = ocb2; osa = osb2;
            rB = rA;
        }
        feed.finish();
        // ---- epilogue: B(k1 - 1) ----
        {
            const int j = k + 1;
            if (STEP != STEP_NONE) {
                ipc_sincos(nb.t, &nsb, &ncb);
                double* pq = SA::stepw(M, j, k - k0, ts.tid);
                pq[0] = nb.x; pq[CS] = nb.y; pq[2 * CS] = nb.t; pq[3 * CS] = ncb; pq[4 * CS] = nsb;
            }
            Lin2 e; double t[NPRE];
            odom_terms<UNI>(O, rB, nca, nsa, na, nb, e, t);
            chi += e.chi; mx = fmax(mx, e.chi);
#pragma unroll
            for (int m = 0; m < NPRE; ++m) run[m] += t[m];
            if (has_spec) {
#pragma unroll
                for (int q = 1; q < NSPEC; ++q) {
                    if (j == spec_v[q]) {
                        double* o = spec + q * SPECW;
#pragma unroll
                        for (int m = 0; m < NPRE; ++m) o[m] = run[m];
                        o[NPRE] = nb.x; o[NPRE + 1] = nb.y; o[NPRE + 2] = nb.t;
                    }
                }
            }
        }
    }
    double s[3] = {chi, hh, gain};
    IPC_PH(1);
    ScanSumMax<NT, 3>::run(run, s, mx, M.red() + (size_t)buf * RED_DOUBLES);
#pragma unroll
    for (int m = 0; m < NPRE; ++m) ts.base[m] = run[m];
    out.chi = s[0]; out.hh = s[1]; out.gain = s[2]; out.mx = mx;
    if (has_spec) {
#pragma unroll
        for (int q = 1; q < NSPEC; ++q) {
            const int v = spec_v[q];
            if (v > k0 && v <= k1) {
                double* o = spec + q * SPECW;
#pragma unroll
                for (int m = 0; m < NPRE; ++m) o[m] += run[m];
            }
        }
    }
    bsync<NT>();
    buf ^= 1;
    IPC_PH(2);
}
template <int NT, bool UNI, bool STG = false, bool GST = false> IPC_HD void sweep(const ChainMem& M, const OdomView& O, int mode, double c1, double c2, ThreadState& ts,
                                              SweepOut& out, int& buf, const int* spec_v IPC_PH_ARG) {
    if (!GST) {
        if (mode == STEP_GN) sweep_mode_direct<NT, UNI, STG, STEP_GN, false>(M, O, c1, c2, ts, out, buf, spec_v IPC_PH_PASS);
        else if (mode == STEP_BLEND) sweep_mode_direct<NT, UNI, STG, STEP_BLEND, false>(M, O, c1, c2, ts, out, buf, spec_v IPC_PH_PASS);
        else sweep_mode_direct<NT, UNI, STG, STEP_NONE, false>(M, O, c1, c2, ts, out, buf, spec_v IPC_PH_PASS);
        return;
    }
    if (mode == STEP_GN) sweep_mode<NT, UNI, STG, STEP_GN, GST>(M, O, c1, c2, ts, out, buf, spec_v IPC_PH_PASS);
    else if (mode == STEP_BLEND) sweep_mode<NT, UNI, STG, STEP_BLEND, GST>(M, O, c1, c2, ts, out, buf, spec_v IPC_PH_PASS);
    else sweep_mode<NT, UNI, STG, STEP_NONE, GST>(M, O, c1, c2, ts, out, buf, spec_v IPC_PH_PASS);
}

// undo the last applied sweep: poses from the backup, cos / sin recomputed, boundary registers re-read. Thread bases are NOT
// restored: every caller re-linearises (GN rejection) or only runs blend sweeps (which do not read them) until a step is kept.
template <int NT, bool GST = false> IPC_HD void rollback(const ChainMem& M, ThreadState& ts) {
    using SA = StateAt<NT, GST>;
    constexpr int CS = SA::CS;
    int sl = ts.tid + 1;
    // global-memory state: the rejected trial went to the other buffer, only the boundary registers are re-read below
#pragma unroll 4
    for (int k = ts.k0; k < ts.k1 && !GST; ++k, sl += NT) {
        const int j = k + 1;
        const double* bq = M.B(sl);
        const double t = bq[2];
        double s, c; ipc_sincos(t, &s, &c);
        double* pq = SA::step(M, j, k - ts.k0, ts.tid);
        pq[0] = bq[0]; pq[CS] = bq[1]; pq[2 * CS] = t; pq[3 * CS] = c; pq[4 * CS] = s;
    }
    bsync<NT>();
    if (ts.k0 < ts.k1 && ts.k0 > 0) {
        const double* pq = SA::step(M, ts.k0, ts.S - 1, ts.tid - 1);         // vertex k0 = the last vertex of thread tid - 1
        ts.pa.x = pq[0]; ts.pa.y = pq[CS]; ts.pa.t = pq[2 * CS]; ts.ca = pq[3 * CS]; ts.sa = pq[4 * CS];
    }
    bsync<NT>();
}


struct SpecVals {            // prefix (PM, Pm) and pose at the special vertices rs, re, L (vertex 0: zeros / origin)
    double pre1[NPRE], pre2[NPRE], pre3[NPRE];
    P2 p1, p2, p3;
};
IPC_HD void spec_load(const double* spec, const CheckGeom& g, SpecVals& sv) {
    const double* o1 = spec + 1 * SPECW; const double* o2 = spec + 2 * SPECW; const double* o3 = spec + 3 * SPECW;
    const bool z1 = g.rs == 0;      // nobody owns vertex 0
#pragma unroll
    for (int m = 0; m < NPRE; ++m) { sv.pre1[m] = z1 ? 0.0 : o1[m]; sv.pre2[m] = o2[m]; sv.pre3[m] = o3[m]; }
    sv.p1.x = z1 ? 0.0 : o1[NPRE]; sv.p1.y = z1 ? 0.0 : o1[NPRE + 1]; sv.p1.t = z1 ? 0.0 : o1[NPRE + 2];
    sv.p2.x = o2[NPRE]; sv.p2.y = o2[NPRE + 1]; sv.p2.t = o2[NPRE + 2];
    sv.p3.x = o3[NPRE]; sv.p3.y = o3[NPRE + 1]; sv.p3.t = o3[NPRE + 2];
}
IPC_HD P2 sel_pose(bool c, const P2& a, const P2& b) { P2 r; r.x = c ? a.x : b.x; r.y = c ? a.y : b.y; r.t = c ? a.t : b.t; return r; }

// A loop edge in twist coordinates is one more edge of the cycle: with Q_l = [c_f -s_f y_t; s_f c_f -x_t; 0 0 1] (from-vertex
// heading, to-vertex position) its terms W_l = Q_l V_l Q_l^T and Q_l d_l are what edge_prefix_terms() computes for an
// odometry edge, and d(residual)/d(interval twist) = sigma Q_l^-1 with sigma = +1 when `to` is the later vertex.
struct LoopNow { Lin2 e; double t[NPRE]; double sigma; double xt, yt; const double* D; const double* V; };
IPC_HD void loop_now(const LoopRec2& L, const P2& pf, const P2& pt, LoopNow& o) {
    double s, c; ipc_sincos(pf.t, &s, &c);
    lin2cs(c, s, pf, pt, L.meas[0], L.meas[1], L.meas[2], L.D, o.e);
    edge_prefix_terms(o.e, L.V, pt.x, pt.y, o.t);
    o.sigma = L.to > L.from ? 1.0 : -1.0;
    o.xt = pt.x; o.yt = pt.y; o.D = L.D; o.V = L.V;
}
IPC_HD void loops_eval(const SpecVals& sv, const CheckGeom& g, const LoopRec2& Lc, const LoopRec2& Lm, LoopNow& lc, LoopNow& lm) {
    const P2 org{0, 0, 0};
    {
        const P2 pa = sel_pose(g.c_a_is_rs, sv.p1, org), pb = sel_pose(g.c_b_is_L, sv.p3, sv.p2);
        const bool to_hi = Lc.to > Lc.from;
        loop_now(Lc, sel_pose(to_hi, pa, pb), sel_pose(to_hi, pb, pa), lc);
    }
    if (g.K == 2) {
        const P2 pa = sel_pose(g.m_a_is_rs, sv.p1, org), pb = sel_pose(g.m_b_is_L, sv.p3, sv.p2);
        const bool to_hi = Lm.to > Lm.from;
        loop_now(Lm, sel_pose(to_hi, pa, pb), sel_pose(to_hi, pb, pa), lm);
    }
}

// GN solution of the published linearisation. Unknowns: the force z_l of every loop on its interval. Stationarity gives the
// SPD system  (P_ll' + delta_ll' W_l) z_l' = q_l + sigma_l Q_l d_l  with P_ll' = sum of PM over the common interval and
// q_l = sum of Pm over the interval of l: 3x3 for K = 1, 6x6 for K = 2 (solved by a 3x3 Schur complement).
// After the step the linearised residual of an edge under force z is -V Q^T z, so
//   model = sum_r z_r^T P_r z_r + sum_l z_l^T W_l z_l.
IPC_HD void gn_solve(const SpecVals& sv, const CheckGeom& g, const LoopNow& lc, const LoopNow& lm, StepSpec* sp) {
    // region sums: region 0 = [0, rs), region 1 = [rs, re), region 2 = [re, L)
    double acc[3][NPRE];
#pragma unroll
    for (int m = 0; m < NPRE; ++m) { acc[0][m] = sv.pre1[m]; acc[1][m] = sv.pre2[m] - sv.pre1[m]; acc[2][m] = sv.pre3[m] - sv.pre2[m]; }
    double zc[3], zm[3] = {0, 0, 0};
    if (g.K == 1) {
        double S[6], Si[6], r[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) S[q] = acc[1][q] + lc.t[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) r[q] = acc[1][6 + q] - lc.sigma * lc.t[6 + q];
        inv_sym3(S, Si);
        sym3_mul(Si, r, zc);
    } else {
        double A[6], B[6], Cm[6], rc[3], rm[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            B[q] = acc[1][q];
            A[q] = acc[1][q] + (g.first_is_c ? acc[0][q] : 0.0) + (g.last_is_c ? acc[2][q] : 0.0) + lc.t[q];
            Cm[q] = acc[1][q] + (g.first_is_c ? 0.0 : acc[0][q]) + (g.last_is_c ? 0.0 : acc[2][q]) + lm.t[q];
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            rc[q] = acc[1][6 + q] + (g.first_is_c ? acc[0][6 + q] : 0.0) + (g.last_is_c ? acc[2][6 + q] : 0.0) - lc.sigma * lc.t[6 + q];
            rm[q] = acc[1][6 + q] + (g.first_is_c ? 0.0 : acc[0][6 + q]) + (g.last_is_c ? 0.0 : acc[2][6 + q]) - lm.sigma * lm.t[6 + q];
        }
        // [A B; B Cm] [zc; zm] = [rc; rm]:  Sch = Cm - B A^-1 B,  zm = Sch^-1 (rm - B A^-1 rc),  zc = A^-1 (rc - B zm)
        double Ai[6], AiB[9], t[3], u[3];
        inv_sym3(A, Ai);
        sym3_sym3(Ai, B, AiB);                                   // A^-1 B (full)
        double Sch[6];
        // B (A^-1 B): symmetric, upper triangle only
        Sch[0] = Cm[0] - (B[0] * AiB[0] + B[1] * AiB[3] + B[2] * AiB[6]);
        Sch[1] = Cm[1] - (B[0] * AiB[1] + B[1] * AiB[4] + B[2] * AiB[7]);
        Sch[2] = Cm[2] - (B[0] * AiB[2] + B[1] * AiB[5] + B[2] * AiB[8]);
        Sch[3] = Cm[3] - (B[1] * AiB[1] + B[3] * AiB[4] + B[4] * AiB[7]);
        Sch[4] = Cm[4] - (B[1] * AiB[2] + B[3] * AiB[5] + B[4] * AiB[8]);
        Sch[5] = Cm[5] - (B[2] * AiB[2] + B[4] * AiB[5] + B[5] * AiB[8]);
        sym3_mul(Ai, rc, t);                                     // A^-1 rc
        sym3_mul(B, t, u);
        double rs2[3] = {rm[0] - u[0], rm[1] - u[1], rm[2] - u[2]};
        double Schi[6];
        inv_sym3(Sch, Schi);
        sym3_mul(Schi, rs2, zm);
        sym3_mul(B, zm, u);
        double rc2[3] = {rc[0] - u[0], rc[1] - u[1], rc[2] - u[2]};
        sym3_mul(Ai, rc2, zc);
    }
    sp->rs = g.rs; sp->re = g.re;
    double z[3][3], C[3][3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (g.K == 1) { z[0][q] = 0; z[1][q] = zc[q]; z[2][q] = 0; }
        else {
            z[0][q] = g.first_is_c ? zc[q] : zm[q];
            z[1][q] = zc[q] + zm[q];
            z[2][q] = g.last_is_c ? zc[q] : zm[q];
        }
    }
    // C_0 = 0, C_1 = PM(rs) (z_0 - z_1), C_2 = C_1 + PM(re) (z_1 - z_2)
    {
        double dz[3], t[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { C[0][q] = 0; dz[q] = z[0][q] - z[1][q]; }
        sym3_mul(sv.pre1, dz, t);
#pragma unroll
        for (int q = 0; q < 3; ++q) { C[1][q] = t[q]; dz[q] = z[1][q] - z[2][q]; }
        sym3_mul(sv.pre2, dz, t);
#pragma unroll
        for (int q = 0; q < 3; ++q) C[2][q] = C[1][q] + t[q];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) { sp->z[r][q] = z[r][q]; sp->C[r][q] = C[r][q]; }
    double model = quad3(lc.t, zc[0], zc[1], zc[2]);
    if (g.K == 2) model += quad3(lm.t, zm[0], zm[1], zm[2]);
    {   // loop part of the predicted gain: residual change of loop l is sigma V Q^T z_l - d_l
        const Lin2& e = lc.e;
        const double xb = lc.xt, yb = lc.yt;
        const double y0 = e.c * zc[0] + e.s * zc[1], y1 = -e.s * zc[0] + e.c * zc[1], y2 = yb * zc[0] - xb * zc[1] + zc[2];
        const double* V = lc.V; const double sg = lc.sigma;
        const double w0 = sg * (V[0] * y0 + V[1] * y1 + V[2] * y2) - e.d0, w1 = sg * (V[1] * y0 + V[3] * y1 + V[4] * y2) - e.d1,
                     w2 = sg * (V[2] * y0 + V[4] * y1 + V[5] * y2) - e.d2;
        double gl = quad3(lc.D, w0, w1, w2);
        if (g.K == 2) {
            const Lin2& f = lm.e;
            const double xm = lm.xt, ym = lm.yt;
            const double u0 = f.c * zm[0] + f.s * zm[1], u1 = -f.s * zm[0] + f.c * zm[1], u2 = ym * zm[0] - xm * zm[1] + zm[2];
            const double* Vm = lm.V; const double sm = lm.sigma;
            const double v0 = sm * (Vm[0] * u0 + Vm[1] * u1 + Vm[2] * u2) - f.d0, v1 = sm * (Vm[1] * u0 + Vm[3] * u1 + Vm[4] * u2) - f.d1,
                         v2 = sm * (Vm[2] * u0 + Vm[4] * u1 + Vm[5] * u2) - f.d2;
            gl += quad3(lm.D, v0, v1, v2);
        }
        sp->gain_loops = gl;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) model += quad3(acc[r], z[r][0], z[r][1], z[r][2]);
    sp->model = model;
}

// After a sweep: thread 0 evaluates the loop edges at the published state and, if the trial is going to be kept
// (rho > 0, or `force`), solves the new linearisation into M.U()->sol. One barrier; every thread gets the loop chi2.
// After a sweep, warp 0: (phase 1) lane l linearises loop l at the published state and stages its terms in shared memory;
// (phase 2) lane 0 decides whether the trial is kept (rho > 0, or `force`) and if so solves the new linearisation into
// M.U()->sol, reading the interval sums straight from the special-vertex table. Everything is staged through shared memory
// so this serial, latency-critical section keeps almost nothing in registers / local memory. (Round 2 A/B: spreading the 6x6 system
// over six lanes — one row per lane, Gaussian elimination with shuffle broadcasts, as the SE(3) kernel does for its 12x12 system —
// LOSES 18 % here (207 K -> 170 K checks/s): twelve fp64 divisions and shuffle round trips are slower than the two reciprocals of the
// 3x3 block-Schur form that one lane runs with full ILP. profiles/r02_ab_log.txt.)
IPC_HD_COLD void eval_and_solve_w0(ChainMem M, int buf, double odom_chi, double cur_chi, double linearGain, bool force) {
    UniBlock* U = M.U();
    const CheckGeom& g = U->g;
    const double* spec = M.spec() + (size_t)(buf ^ 1) * NSPEC * SPECW;
    const double* o1 = spec + 1 * SPECW; const double* o2 = spec + 2 * SPECW; const double* o3 = spec + 3 * SPECW;
    const bool z1 = g.rs == 0;      // nobody owns vertex 0: prefix 0, pose = origin
#ifdef __CUDA_ARCH__
    const int l0 = threadIdx.x, l1 = threadIdx.x + 1;
#else
    const int l0 = 0, l1 = g.K;
#endif
    for (int l = l0; l < l1 && l < g.K; ++l) {
        const LoopRec2& Lp = l == 0 ? U->Lc : U->Lm;
        const bool a_is_rs = l == 0 ? g.c_a_is_rs : g.m_a_is_rs, b_is_L = l == 0 ? g.c_b_is_L : g.m_b_is_L;
        // interval [a, b): a is vertex 0 or rs, b is re or L
        const double* oa = o1; const double* ob = b_is_L ? o3 : o2;
        const bool a_org = !a_is_rs || z1;
        const P2 pa{a_org ? 0.0 : oa[NPRE], a_org ? 0.0 : oa[NPRE + 1], a_org ? 0.0 : oa[NPRE + 2]};
        const P2 pb{ob[NPRE], ob[NPRE + 1], ob[NPRE + 2]};
        const bool to_hi = Lp.to > Lp.from;
        const P2 pf = to_hi ? pa : pb, pt = to_hi ? pb : pa;
        double sn, cs; ipc_sincos(pf.t, &sn, &cs);
        Lin2 e; lin2cs(cs, sn, pf, pt, Lp.meas[0], Lp.meas[1], Lp.meas[2], Lp.D, e);
        double* o = U->lt[l];
        edge_prefix_terms(e, Lp.V, pt.x, pt.y, o);
        o[9] = to_hi ? 1.0 : -1.0; o[10] = e.chi; o[11] = e.d0; o[12] = e.d1; o[13] = e.d2; o[14] = e.c; o[15] = e.s; o[16] = pt.x; o[17] = pt.y;
    }
#ifdef __CUDA_ARCH__
    __syncwarp();
    if (threadIdx.x != 0) return;
#endif
    const double* tc = U->lt[0]; const double* tm = U->lt[1];
    const double c = tc[10], m = g.K == 2 ? tm[10] : 0.0;
    if (fabs(linearGain) < 1e-12) linearGain = 1e-12;
    const double rho = (cur_chi - (odom_chi + c + m)) / linearGain;
    U->n_c = c; U->n_m = m;
    if (!(force || rho > 0)) return;
    StepSpec* sp = &U->sol;
    // region sums: region 0 = [0, rs), region 1 = [rs, re), region 2 = [re, L), read on the fly:
    //   acc0 = pre1, acc1 = pre2 - pre1, acc2 = pre3 - pre2  (pre1 = 0 when rs == 0)
    auto P1 = [&](int q) { return z1 ? 0.0 : o1[q]; };
    double zc[3], zm[3] = {0, 0, 0};
    if (g.K == 1) {
        double S[6], Si[6], r[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) S[q] = (o2[q] - P1(q)) + tc[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) r[q] = (o2[6 + q] - P1(6 + q)) - tc[9] * tc[6 + q];
        inv_sym3(S, Si);
        sym3_mul(Si, r, zc);
    } else {
        double A[6], B[6], Cm[6], rc[3], rm[3];
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const double a0 = P1(q), a1 = o2[q] - a0, a2 = o3[q] - o2[q];
            const double cc = a1 + (g.first_is_c ? a0 : 0.0) + (g.last_is_c ? a2 : 0.0);
            const double mm = a1 + (g.first_is_c ? 0.0 : a0) + (g.last_is_c ? 0.0 : a2);
            if (q < 6) { B[q] = a1; A[q] = cc + tc[q]; Cm[q] = mm + tm[q]; }
            else { rc[q - 6] = cc - tc[9] * tc[q]; rm[q - 6] = mm - tm[9] * tm[q]; }
        }
        // [A B; B Cm] [zc; zm] = [rc; rm]:  Sch = Cm - B A^-1 B,  zm = Sch^-1 (rm - B A^-1 rc),  zc = A^-1 (rc - B zm)
        double Ai[6], AiB[9], t[3], u[3];
        inv_sym3(A, Ai);
        sym3_sym3(Ai, B, AiB);
        double Sch[6];
        Sch[0] = Cm[0] - (B[0] * AiB[0] + B[1] * AiB[3] + B[2] * AiB[6]);
        Sch[1] = Cm[1] - (B[0] * AiB[1] + B[1] * AiB[4] + B[2] * AiB[7]);
        Sch[2] = Cm[2] - (B[0] * AiB[2] + B[1] * AiB[5] + B[2] * AiB[8]);
        Sch[3] = Cm[3] - (B[1] * AiB[1] + B[3] * AiB[4] + B[4] * AiB[7]);
        Sch[4] = Cm[4] - (B[1] * AiB[2] + B[3] * AiB[5] + B[4] * AiB[8]);
        Sch[5] = Cm[5] - (B[2] * AiB[2] + B[4] * AiB[5] + B[5] * AiB[8]);
        sym3_mul(Ai, rc, t);
        sym3_mul(B, t, u);
        const double rs2[3] = {rm[0] - u[0], rm[1] - u[1], rm[2] - u[2]};
        double Schi[6];
        inv_sym3(Sch, Schi);
        sym3_mul(Schi, rs2, zm);
        sym3_mul(B, zm, u);
        const double rc2[3] = {rc[0] - u[0], rc[1] - u[1], rc[2] - u[2]};
        sym3_mul(Ai, rc2, zc);
    }
    sp->rs = g.rs; sp->re = g.re;
    double z0[3], zz1[3], z2[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (g.K == 1) { z0[q] = 0; zz1[q] = zc[q]; z2[q] = 0; }
        else { z0[q] = g.first_is_c ? zc[q] : zm[q]; zz1[q] = zc[q] + zm[q]; z2[q] = g.last_is_c ? zc[q] : zm[q]; }
        sp->z[0][q] = z0[q]; sp->z[1][q] = zz1[q]; sp->z[2][q] = z2[q];
    }
    {   // C_0 = 0, C_1 = PM(rs) (z_0 - z_1), C_2 = C_1 + PM(re) (z_1 - z_2)
        const double dz0[3] = {z0[0] - zz1[0], z0[1] - zz1[1], z0[2] - zz1[2]}, dz1[3] = {zz1[0] - z2[0], zz1[1] - z2[1], zz1[2] - z2[2]};
        const double p1[6] = {P1(0), P1(1), P1(2), P1(3), P1(4), P1(5)};
        double c1v[3], c2v[3];
        sym3_mul(p1, dz0, c1v);
        sym3_mul(o2, dz1, c2v);
#pragma unroll
        for (int q = 0; q < 3; ++q) { sp->C[0][q] = 0; sp->C[1][q] = c1v[q]; sp->C[2][q] = c1v[q] + c2v[q]; }
    }
    {   // the loop part of the predicted gain. (The model value chi2 - model, five more quadratic forms, is not evaluated here any more:
        // nothing reads it in the SE(2) kernels since the gain is accumulated edge by edge, and this section is on the critical path of
        // every sweep.)
        sp->model = 0.0;
        double gl = 0;
        for (int l = 0; l < g.K; ++l) {   // residual change of loop l is sigma V Q^T z_l - d_l
            const double* o = U->lt[l];
            const LoopRec2& Lp = l == 0 ? U->Lc : U->Lm;
            const double* zl = l == 0 ? zc : zm;
            const double y0 = o[14] * zl[0] + o[15] * zl[1], y1 = -o[15] * zl[0] + o[14] * zl[1], y2 = o[17] * zl[0] - o[16] * zl[1] + zl[2];
            const double* V = Lp.V; const double sg = o[9];
            const double w0 = sg * (V[0] * y0 + V[1] * y1 + V[2] * y2) - o[11], w1 = sg * (V[1] * y0 + V[3] * y1 + V[4] * y2) - o[12],
                         w2 = sg * (V[2] * y0 + V[4] * y1 + V[5] * y2) - o[13];
            gl += quad3(Lp.D, w0, w1, w2);
        }
        sp->gain_loops = gl;
    }
}
template <int NT> IPC_HD void eval_and_solve(const ChainMem& M, int buf, double odom_chi, double cur_chi, double linearGain, bool force, double& n_c,
                                             double& n_m) {
#ifdef __CUDA_ARCH__
    if (threadIdx.x < 32) eval_and_solve_w0(M, buf, odom_chi, cur_chi, linearGain, force);
#else
    if (hd_tid() == 0) eval_and_solve_w0(M, buf, odom_chi, cur_chi, linearGain, force);
#endif
    bsync<NT>();
    n_c = M.U()->n_c; n_m = M.U()->n_m;
}

// |h_gn|^2 of the current linearisation without applying it (state in shared memory: read in place)
template <int NT, bool UNI, bool STG = false, bool GST = false> IPC_HD double gn_norm_sq_direct(const ChainMem& M, const OdomView& O, const ThreadState& ts) {
    using SA = StateAt<NT, GST>;
    constexpr int CS = SA::CS;
    double v[1] = {0};
    const StepSpec* sp = &M.U()->sol;
    double pre[NPRE];
#pragma unroll
    for (int m = 0; m < NPRE; ++m) pre[m] = ts.base[m];
    P2 pa = ts.pa; double ca = ts.ca, sa = ts.sa;
    OdomRec<UNI> rn;
    int es = ts.tid;
    if (ts.k0 < ts.k1) odom_load<UNI, STG>(O, es, ts.k0, rn);
    for (int k = ts.k0; k < ts.k1; ++k, es += NT) {
        const int j = k + 1;
        const double* pq = SA::step(M, j, k - ts.k0, ts.tid);
        const P2 pb{pq[0], pq[CS], pq[2 * CS]};
        const double cb = pq[3 * CS], sb = pq[4 * CS];
        const OdomRec<UNI> r = rn;
        if (j < ts.k1) odom_load<UNI, STG>(O, es + NT, j, rn);
        Lin2 e; double t[NPRE], h[3];
        odom_terms<UNI>(O, r, ca, sa, pa, pb, e, t);
#pragma unroll
        for (int m = 0; m < NPRE; ++m) pre[m] += t[m];
        gn_step_at(sp, j, pre, pb.x, pb.y, h);
        v[0] += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
        pa = pb; ca = cb; sa = sb;
    }
    hd_block_sum<NT, 1>(v, M.red2());
    return v[0];
}

// |h_gn|^2 of the current linearisation without applying it
template <int NT, bool UNI, bool STG = false, bool GST = false> IPC_HD double gn_norm_sq(const ChainMem& M, const OdomView& O, const ThreadState& ts) {
    if (!GST) return gn_norm_sq_direct<NT, UNI, STG, false>(M, O, ts);
    double v[1] = {0};
    const StepSpec* sp = &M.U()->sol;
    double pre[NPRE];
#pragma unroll
    for (int m = 0; m < NPRE; ++m) pre[m] = ts.base[m];
    P2 pa = ts.pa; double ca = ts.ca, sa = ts.sa;
    if (ts.k0 < ts.k1) {
        TileFeed<NT, UNI, STG, GST, false> feed(M, O, ts.k0, ts.k1 - ts.k0, ts.tid);
        feed.start();
        for (int k = ts.k0; k < ts.k1; ++k) {
            const int j = k + 1;
            double p5[5], g6[6]; OdomRec<UNI> r;
            feed.next(k - ts.k0, p5, r, g6);
            const P2 pb{p5[0], p5[1], p5[2]};
            Lin2 e; double t[NPRE], h[3];
            odom_terms<UNI>(O, r, ca, sa, pa, pb, e, t);
#pragma unroll
            for (int m = 0; m < NPRE; ++m) pre[m] += t[m];
            gn_step_at(sp, j, pre, pb.x, pb.y, h);
            v[0] += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
            pa = pb; ca = p5[3]; sa = p5[4];
        }
        feed.finish();
    }
    hd_block_sum<NT, 1>(v, M.red2());
    return v[0];
}

// Steepest-descent pass at the current state (poses + prefixes valid): gradient b and h_gn per vertex (g2o vertex coordinates)
// into the scratch, bb = |b|^2, bh = b . h_gn, hh = |h_gn|^2, bHb = b^T H b = sum over edges |J (b_k, b_k+1)|^2_D — in ONE pass:
// the edge term of b^T H b lags one edge behind the gradient, so both are finished from a single evaluation of every edge. A thread owns the vertices
// k0+1..k1 and the Hessian terms of the edges k0+1..k1 (thread 0 also edge 0); the one term that needs the next thread's first
// gradient is completed after the block barrier from the scratch. Stands in for gn_norm_sq + a separate gradient pass when the
// iteration is expected to be trust-region bound (gn_norm_sq alone is the cheaper pass when the GN step is expected to fit).
template <int NT, bool UNI, bool STG = false, bool GST = false> IPC_HD void sd_fused_direct(const ChainMem& M, const OdomView& O, const ThreadState& ts, double& bb,
                                                 double& bh, double& hh, double& bHb) {
    using SA = StateAt<NT, GST>;
    constexpr int CS = SA::CS;
    const CheckGeom& g = M.U()->g;
    const int k0 = ts.k0, k1 = ts.k1, L = g.L;
    const StepSpec* sp = &M.U()->sol;
    const LoopRec2& Lc = M.U()->Lc; const LoopRec2& Lm = M.U()->Lm;
    Lin2 ec, em; const int cjf = Lc.from - g.lo, cjt = Lc.to - g.lo; int mjf = -1, mjt = -1;
    {
        const double* qf = SA::vertex(M, cjf, ts.S); const double* qt = SA::vertex(M, cjt, ts.S);
        P2 pf{qf[0], qf[CS], qf[2 * CS]}, pt{qt[0], qt[CS], qt[2 * CS]};
        lin2cs(qf[3 * CS], qf[4 * CS], pf, pt, Lc.meas[0], Lc.meas[1], Lc.meas[2], Lc.D, ec);
    }
    double gci[3], gcj[3], gmi[3] = {0, 0, 0}, gmj[3] = {0, 0, 0};
    grad2(ec, gci, gcj);
    if (g.K == 2) {
        mjf = Lm.from - g.lo; mjt = Lm.to - g.lo;
        const double* qf = SA::vertex(M, mjf, ts.S); const double* qt = SA::vertex(M, mjt, ts.S);
        P2 pf{qf[0], qf[CS], qf[2 * CS]}, pt{qt[0], qt[CS], qt[2 * CS]};
        lin2cs(qf[3 * CS], qf[4 * CS], pf, pt, Lm.meas[0], Lm.meas[1], Lm.meas[2], Lm.D, em);
        grad2(em, gmi, gmj);
    }
    double v[4] = {0, 0, 0, 0};
    double pre[NPRE];
#pragma unroll
    for (int m = 0; m < NPRE; ++m) pre[m] = ts.base[m];
    double bprev[3] = {0, 0, 0};        // gradient at the tail vertex of the lagging edge (vertex 0: fixed, zero)
    double pc = 1, ps = 0, prx = 0, pry = 0, pD[6] = {0, 0, 0, 0, 0, 0};   // Jacobian pieces (and information) of the lagging edge
    bool tail_pending = false;          // edge k1 (< L) waits for the next thread's first gradient
    auto finish_vertex = [&](int j, int sl, const double* gsum, double x, double y, double* b) {
        b[0] = -gsum[0]; b[1] = -gsum[1]; b[2] = -gsum[2];
        if (j == cjf) { b[0] -= gci[0]; b[1] -= gci[1]; b[2] -= gci[2]; }
        if (j == cjt) { b[0] -= gcj[0]; b[1] -= gcj[1]; b[2] -= gcj[2]; }
        if (j == mjf) { b[0] -= gmi[0]; b[1] -= gmi[1]; b[2] -= gmi[2]; }
        if (j == mjt) { b[0] -= gmj[0]; b[1] -= gmj[1]; b[2] -= gmj[2]; }
        double h[3];
        gn_step_at(sp, j, pre, x, y, h);
#pragma unroll
        for (int q = 0; q < 3; ++q) { M.G(sl)[q] = b[q]; M.G(sl)[3 + q] = h[q]; }
        v[0] += b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
        v[1] += b[0] * h[0] + b[1] * h[1] + b[2] * h[2];
        v[2] += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
    };
    auto lag_term = [&](const double* ba, const double* bv) {   // |J (ba, bv)|^2_D of the lagging edge
        const double ux = bv[0] - ba[0], uy = bv[1] - ba[1];
        const double q0 = pc * ux + ps * uy + pry * ba[2], q1 = -ps * ux + pc * uy - prx * ba[2], q2 = bv[2] - ba[2];
        v[3] += quad3(UNI ? O.Du : pD, q0, q1, q2);
    };
    if (k0 < k1) {
        P2 pa = ts.pa; double ca = ts.ca, sa = ts.sa;
        double gprev[3] = {0, 0, 0};
        int sl = ts.tid + 1;
        int es = ts.tid;
        OdomRec<UNI> rn; odom_load<UNI, STG>(O, es, k0, rn);
        for (int k = k0; k <= k1 && k < L; ++k, es += NT) {
            const double* pq = k < k1 ? SA::step(M, k + 1, k - k0, ts.tid) : SA::step(M, k + 1, 0, ts.tid + 1);   // vertex k1 + 1: first of thread tid + 1
            P2 pb{pq[0], pq[CS], pq[2 * CS]};
            const double cbn = pq[3 * CS], sbn = pq[4 * CS];
            Lin2 e; double t[NPRE];
            const OdomRec<UNI> r = rn;
            if (k + 1 <= k1 && k + 1 < L) odom_load<UNI, STG>(O, k + 1 < k1 ? es + NT : ts.tid + 1, k + 1, rn);   // edge k1: first of thread tid + 1
            odom_terms<UNI>(O, r, ca, sa, pa, pb, e, t);
            double gi[3], gj[3]; grad2(e, gi, gj);
            if (k > k0) {
                const double gs[3] = {gprev[0] + gi[0], gprev[1] + gi[1], gprev[2] + gi[2]};
                double b[3];
                finish_vertex(k, sl, gs, pa.x, pa.y, b);     // vertex k = k0 + 1 + (k - k0 - 1)
                sl += NT;
                if (k - 1 > k0 || k0 == 0) lag_term(bprev, b);      // edge k-1: both gradients are this thread's
                bprev[0] = b[0]; bprev[1] = b[1]; bprev[2] = b[2];
            }
#pragma unroll
            for (int m = 0; m < NPRE; ++m) pre[m] += t[m];
            gprev[0] = gj[0]; gprev[1] = gj[1]; gprev[2] = gj[2];
            pc = e.c; ps = e.s; prx = e.rx; pry = e.ry;
            if (!UNI) {
#pragma unroll
                for (int c = 0; c < 6; ++c) pD[c] = r.z[UNI ? 0 : 3 + c];
            }
            pa = pb; ca = cbn; sa = sbn;
        }
        if (k1 == L) {
            double b[3];
            finish_vertex(L, sl, gprev, pa.x, pa.y, b);
            if (L - 1 > k0 || k0 == 0) lag_term(bprev, b);
        } else tail_pending = true;
    }
    if (hd_tid() == 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) { M.G(0)[q] = 0; M.G(0)[3 + q] = 0; }
    }
    bsync<NT>();                        // every gradient is in the scratch
    if (tail_pending) {
        const double* gq = M.G(ts.tid + 2);         // vertex k1 + 1 = the first vertex of thread tid + 1
        const double bn[3] = {gq[0], gq[1], gq[2]};
        lag_term(bprev, bn);
    }
    if (hd_tid() == 0) {
        const double* gf = M.G(vslot<NT>(cjf, ts.S)); const double* gt = M.G(vslot<NT>(cjt, ts.S));
        double bf[3] = {gf[0], gf[1], gf[2]}, bt[3] = {gt[0], gt[1], gt[2]}, q0, q1, q2;
        dlin2(ec, bf, bt, q0, q1, q2); v[3] += quad3(Lc.D, q0, q1, q2);
        if (g.K == 2) {
            gf = M.G(vslot<NT>(mjf, ts.S)); gt = M.G(vslot<NT>(mjt, ts.S));
            double mf[3] = {gf[0], gf[1], gf[2]}, mt[3] = {gt[0], gt[1], gt[2]};
            dlin2(em, mf, mt, q0, q1, q2); v[3] += quad3(Lm.D, q0, q1, q2);
        }
    }
    hd_block_sum<NT, 4>(v, M.red2());
    bb = v[0]; bh = v[1]; hh = v[2]; bHb = v[3];
}

// Steepest-descent pass at the current state (poses + prefixes valid): gradient b and h_gn per vertex (g2o vertex coordinates)
// into the scratch, bb = |b|^2, bh = b . h_gn, hh = |h_gn|^2, bHb = b^T H b = sum over edges |J (b_k, b_k+1)|^2_D — in ONE pass:
// the edge term of b^T H b lags one edge behind the gradient, so both are finished from a single evaluation of every edge. A thread owns the vertices
// k0+1..k1 and the Hessian terms of the edges k0+1..k1 (thread 0 also edge 0); the one term that needs the next thread's first
// gradient is completed after the block barrier from the scratch. Stands in for gn_norm_sq + a separate gradient pass when the
// iteration is expected to be trust-region bound (gn_norm_sq alone is the cheaper pass when the GN step is expected to fit).
template <int NT, bool UNI, bool STG = false, bool GST = false> IPC_HD void sd_fused(const ChainMem& M, const OdomView& O, const ThreadState& ts, double& bb,
                                                 double& bh, double& hh, double& bHb) {
    if (!GST) { sd_fused_direct<NT, UNI, STG, false>(M, O, ts, bb, bh, hh, bHb); return; }
    using SA = StateAt<NT, GST>;
    constexpr int CS = SA::CS;
    const CheckGeom& g = M.U()->g;
    const int k0 = ts.k0, k1 = ts.k1, L = g.L;
    const StepSpec* sp = &M.U()->sol;
    const LoopRec2& Lc = M.U()->Lc; const LoopRec2& Lm = M.U()->Lm;
    Lin2 ec, em; const int cjf = Lc.from - g.lo, cjt = Lc.to - g.lo; int mjf = -1, mjt = -1;
    {
        const double* qf = SA::vertex(M, cjf, ts.S); const double* qt = SA::vertex(M, cjt, ts.S);
        P2 pf{qf[0], qf[CS], qf[2 * CS]}, pt{qt[0], qt[CS], qt[2 * CS]};
        lin2cs(qf[3 * CS], qf[4 * CS], pf, pt, Lc.meas[0], Lc.meas[1], Lc.meas[2], Lc.D, ec);
    }
    double gci[3], gcj[3], gmi[3] = {0, 0, 0}, gmj[3] = {0, 0, 0};
    grad2(ec, gci, gcj);
    if (g.K == 2) {
        mjf = Lm.from - g.lo; mjt = Lm.to - g.lo;
        const double* qf = SA::vertex(M, mjf, ts.S); const double* qt = SA::vertex(M, mjt, ts.S);
        P2 pf{qf[0], qf[CS], qf[2 * CS]}, pt{qt[0], qt[CS], qt[2 * CS]};
        lin2cs(qf[3 * CS], qf[4 * CS], pf, pt, Lm.meas[0], Lm.meas[1], Lm.meas[2], Lm.D, em);
        grad2(em, gmi, gmj);
    }
    double v[4] = {0, 0, 0, 0};
    double pre[NPRE];
#pragma unroll
    for (int m = 0; m < NPRE; ++m) pre[m] = ts.base[m];
    double bprev[3] = {0, 0, 0};        // gradient at the tail vertex of the lagging edge (vertex 0: fixed, zero)
    double pc = 1, ps = 0, prx = 0, pry = 0, pD[6] = {0, 0, 0, 0, 0, 0};   // Jacobian pieces (and information) of the lagging edge
    bool tail_pending = false;          // edge k1 (< L) waits for the next thread's first gradient
    auto finish_vertex = [&](int j, int sl, const double* gsum, double x, double y, double* b) {
        b[0] = -gsum[0]; b[1] = -gsum[1]; b[2] = -gsum[2];
        if (j == cjf) { b[0] -= gci[0]; b[1] -= gci[1]; b[2] -= gci[2]; }
        if (j == cjt) { b[0] -= gcj[0]; b[1] -= gcj[1]; b[2] -= gcj[2]; }
        if (j == mjf) { b[0] -= gmi[0]; b[1] -= gmi[1]; b[2] -= gmi[2]; }
        if (j == mjt) { b[0] -= gmj[0]; b[1] -= gmj[1]; b[2] -= gmj[2]; }
        double h[3];
        gn_step_at(sp, j, pre, x, y, h);
#pragma unroll
        for (int q = 0; q < 3; ++q) { M.G(sl)[q] = b[q]; M.G(sl)[3 + q] = h[q]; }
        v[0] += b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
        v[1] += b[0] * h[0] + b[1] * h[1] + b[2] * h[2];
        v[2] += h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
    };
    auto lag_term = [&](const double* ba, const double* bv) {   // |J (ba, bv)|^2_D of the lagging edge
        const double ux = bv[0] - ba[0], uy = bv[1] - ba[1];
        const double q0 = pc * ux + ps * uy + pry * ba[2], q1 = -ps * ux + pc * uy - prx * ba[2], q2 = bv[2] - ba[2];
        v[3] += quad3(UNI ? O.Du : pD, q0, q1, q2);
    };
    if (k0 < k1) {
        P2 pa = ts.pa; double ca = ts.ca, sa = ts.sa;
        double gprev[3] = {0, 0, 0};
        int sl = ts.tid + 1;
        TileFeed<NT, UNI, STG, GST, false> feed(M, O, k0, k1 - k0, ts.tid);
        feed.start();
        for (int k = k0; k <= k1 && k < L; ++k) {
            double p5[5], g6u[6]; OdomRec<UNI> r;
            if (k < k1) feed.next(k - k0, p5, r, g6u);
            else {                               // edge k1: the first step of thread tid + 1 (vertex k1 + 1, its record)
                const double* pq = SA::step(M, k + 1, 0, ts.tid + 1);
#pragma unroll
                for (int q = 0; q < 5; ++q) p5[q] = pq[q * CS];
                if (GST && UNI && !STG && (IPC_RING_DEVICE != 0) && (IPC_RING_ODOM_SHARED != 0)) {      // no per-check copy of the records
#pragma unroll
                    for (int q = 0; q < 3; ++q) r.z[q] = ldg_d(odom_rec<UNI>(O, k) + q);
                } else odom_load<UNI, STG>(O, ts.tid + 1, k, r);
            }
            P2 pb{p5[0], p5[1], p5[2]};
            const double cbn = p5[3], sbn = p5[4];
            Lin2 e; double t[NPRE];
            odom_terms<UNI>(O, r, ca, sa, pa, pb, e, t);
            double gi[3], gj[3]; grad2(e, gi, gj);
            if (k > k0) {
                const double gs[3] = {gprev[0] + gi[0], gprev[1] + gi[1], gprev[2] + gi[2]};
                double b[3];
                finish_vertex(k, sl, gs, pa.x, pa.y, b);     // vertex k = k0 + 1 + (k - k0 - 1)
                sl += NT;
                if (k - 1 > k0 || k0 == 0) lag_term(bprev, b);      // edge k-1: both gradients are this thread's
                bprev[0] = b[0]; bprev[1] = b[1]; bprev[2] = b[2];
            }
#pragma unroll
            for (int m = 0; m < NPRE; ++m) pre[m] += t[m];
            gprev[0] = gj[0]; gprev[1] = gj[1]; gprev[2] = gj[2];
            pc = e.c; ps = e.s; prx = e.rx; pry = e.ry;
            if (!UNI) {
#pragma unroll
                for (int c = 0; c < 6; ++c) pD[c] = r.z[UNI ? 0 : 3 + c];
            }
            pa = pb; ca = cbn; sa = sbn;
        }
        feed.finish();
        if (k1 == L) {
            double b[3];
            finish_vertex(L, sl, gprev, pa.x, pa.y, b);
            if (L - 1 > k0 || k0 == 0) lag_term(bprev, b);
        } else tail_pending = true;
    }
    if (hd_tid() == 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) { M.G(0)[q] = 0; M.G(0)[3 + q] = 0; }
    }
    bsync<NT>();                        // every gradient is in the scratch
    if (tail_pending) {
        const double* gq = M.G(ts.tid + 2);         // vertex k1 + 1 = the first vertex of thread tid + 1
        const double bn[3] = {gq[0], gq[1], gq[2]};
        lag_term(bprev, bn);
    }
    if (hd_tid() == 0) {
        const double* gf = M.G(vslot<NT>(cjf, ts.S)); const double* gt = M.G(vslot<NT>(cjt, ts.S));
        double bf[3] = {gf[0], gf[1], gf[2]}, bt[3] = {gt[0], gt[1], gt[2]}, q0, q1, q2;
        dlin2(ec, bf, bt, q0, q1, q2); v[3] += quad3(Lc.D, q0, q1, q2);
        if (g.K == 2) {
            gf = M.G(vslot<NT>(mjf, ts.S)); gt = M.G(vslot<NT>(mjt, ts.S));
            double mf[3] = {gf[0], gf[1], gf[2]}, mt[3] = {gt[0], gt[1], gt[2]};
            dlin2(em, mf, mt, q0, q1, q2); v[3] += quad3(Lm.D, q0, q1, q2);
        }
    }
    hd_block_sum<NT, 4>(v, M.red2());
    bb = v[0]; bh = v[1]; hh = v[2]; bHb = v[3];
}

struct CheckParams {
    double fast_th, slow_th;
    int fast_iter, slow_iter;
    double noise_eps;        // > 0: stop retrying once a rejected trial's own predicted gain is <= noise_eps * chi2 (below the
                             // round-off of the chi2 evaluation every later, smaller, retry is a coin flip on noise); 0 = replay all retries
    int max_tries, speculate, early_accept;
    int sd_fuse;             // when the GN norm is needed before a trial: 0 = norm pass, then the steepest-descent pass if the step
                             // does not fit; 1 = always the steepest-descent pass (it also delivers the norm); 2 (default) = the
                             // steepest-descent pass when the previous decision was trust-region bound, else the norm pass
};
struct CheckResult {
    int verdict;
    double max_chi2, cand_chi2, sum_chi2;
    int iterations, evals, window_len, n_loops;
    int n_sweeps;            // diagnostics: passes over the chain of every kind (trial / re-linearisation sweeps, norm and gradient passes)
    int n_norm, n_sd, n_relin, n_blend;   // diagnostics: norm pre-passes, steepest-descent pass pairs, re-linearisation sweeps, blend trials
};

// One check, executed cooperatively by NT threads (device) or by the calling thread (host, NT = 1).
// The Dogleg of g2o (OptimizationAlgorithmDogleg::solve inside SparseOptimizer::optimize) is written as a state machine
// around ONE sweep call site, so the sweep is inlined exactly once and nothing lives in local memory.
enum { P_INIT = 0, P_TRIAL_SPEC, P_TRIAL_GN, P_TRIAL_BLEND, P_RELIN_SPECFAIL, P_RELIN_REJECT };

// STG (device, UNI only): `stage` = shared-memory buffer for the window's odometry records + an mbarrier behind it (StageMem).
struct StageMem { double* buf; unsigned long long* mbar; unsigned phase; };
template <int NT, bool UNI, bool STG = false, bool GST = false> IPC_HD void run_check(const ChainMem& M_in, const double* odom, const double* Du, const double* Vu,
                                                  const LoopRec2* Lc_in, const LoopRec2* Lm_in, const CheckParams& prm, bool want_info, CheckResult& res,
                                                  StageMem* stage = nullptr) {
    using SA = StateAt<NT, GST>;
    constexpr int CS = SA::CS;
    ChainMem M = M_in;                                      // st / stw swap when a trial is accepted (global-memory state)
    const int tid = hd_tid();
#if defined(IPC_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
    const long long ph_start = clock64();
#endif
    bsync<NT>();                                            // previous check is done with every array and with the UniBlock
    if (tid == 0) {
        UniBlock* U = M.U();
        U->Lc = *Lc_in; U->Lm = Lm_in ? *Lm_in : *Lc_in;
        CheckGeom g;
        const int cf = U->Lc.from, ct = U->Lc.to;
        const int ca = cf < ct ? cf : ct, cb = cf < ct ? ct : cf;
        int lo = ca, hi = cb; g.K = 1;
        int ma = 0, mb = 0;
        if (Lm_in) {
            const int mf = U->Lm.from, mt = U->Lm.to;
            ma = mf < mt ? mf : mt; mb = mf < mt ? mt : mf;
            // src/consensus.cpp:157-159: positive-length overlap pulls the member into the cluster
            if ((mb < cb ? mb : cb) - (ma > ca ? ma : ca) > 0) { g.K = 2; lo = ca < ma ? ca : ma; hi = cb > mb ? cb : mb; }
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
        g.spec_v[0] = 0; g.spec_v[1] = g.rs; g.spec_v[2] = g.re; g.spec_v[3] = L;
        g.c_a_is_rs = ca_l != 0; g.c_b_is_L = cb_l == L; g.m_a_is_rs = ma_l != 0; g.m_b_is_L = mb_l == L;
        U->g = g;
    }
    bsync<NT>();
    const int K = M.U()->g.K, L = M.U()->g.L;
    const int spec_v[NSPEC] = {0, M.U()->g.rs, M.U()->g.re, L};
    const double th = (K == 2) ? prm.slow_th : prm.fast_th;
    int max_iter = (K == 2) ? prm.slow_iter : prm.fast_iter;
    if (L + K > 100) max_iter *= 5;                        // src/consensus_utils.cpp:12-13
    OdomView O{odom + (UNI ? 3 : 9) * (size_t)M.U()->g.lo, Du, Vu, M.Z(), nullptr};
#ifdef __CUDA_ARCH__
    if (STG) {
        // ONE bulk asynchronous copy (TMA engine, 1-D) of the window's records into shared memory. Source and size must be multiples
        // of 16 B: records are 24 B, so the copy starts at the even record at or below lo and is rounded up (the array is padded).
        const int lo = M.U()->g.lo, off = lo & 1;
        const unsigned bytes = (unsigned)((24u * (unsigned)(L + off) + 15u) & ~15u);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(stage->buf), bar = (unsigned)__cvta_generic_to_shared(stage->mbar);
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(odom + 3 * (size_t)(lo - off)), "r"(bytes), "r"(bar)
                         : "memory");
        }
        asm volatile(
            "{\n .reg .pred p;\n IPC_STG_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra IPC_STG_DONE;\n bra IPC_STG_WAIT;\n IPC_STG_DONE:\n }" ::"r"(bar),
            "r"(stage->phase)
            : "memory");
        stage->phase ^= 1u;
        O.sm = stage->buf + 3 * off;
    }
#endif

    ThreadState ts;
    int S = (L + NT - 1) / NT; if (S < 1) S = 1; S |= 1;   // odd segment length: conflict-free strided shared-memory access
    ts.k0 = tid * S < L ? tid * S : L; ts.k1 = ts.k0 + S < L ? ts.k0 + S : L;
    ts.S = S; ts.tid = tid;
    const int k0 = ts.k0, k1 = ts.k1;

    // ---- dead-reckoning (propagateGuess, src/consensus_utils.cpp:98-116) as two block scans --------------
    // The loops walk in chunks of four steps with the loads of a chunk issued together: when the state lives in global memory a
    // step-by-step loop would pay one L2 / HBM round trip per step (loads cannot be hoisted above the state stores by the compiler).
    // ODOM_DIRECT (ring kernels): the records come straight from the graph's array in every pass, no per-check copy is made.
    constexpr bool ODOM_DIRECT = GST && UNI && !STG && (IPC_RING_DEVICE != 0) && (IPC_RING_ODOM_SHARED != 0);
    {
        double v[1] = {0};
        {   // the only strided read of the odometry: the records go to the scratch in slot order
            int es = tid;
            if (STG) { for (int k = k0; k < k1; ++k) v[0] += O.sm[3 * k + 2]; }
            else if (ODOM_DIRECT) {
#pragma unroll 4
                for (int k = k0; k < k1; ++k) v[0] += ldg_d(odom_rec<UNI>(O, k) + 2);
            } else
#pragma unroll 4
            for (int k = k0; k < k1; ++k, es += NT) {
                const double* zr = odom_rec<UNI>(O, k);
                double* zo = O.zs + (UNI ? 3 : 9) * (size_t)es;
#pragma unroll
                for (int q = 0; q < (UNI ? 3 : 9); ++q) { const double x = ldg_d(zr + q); zo[q] = x; if (q == 2) v[0] += x; }
            }
        }
        hd_block_excl_scan<NT, 1>(v, M.red2());
        double acc = v[0];
        const double th0 = wrap_pi_hd(acc);                 // heading of vertex k0
        if (tid == 0) {
            double* p0 = SA::vertex(M, 0, 1); p0[0] = 0; p0[CS] = 0; p0[2 * CS] = 0; p0[3 * CS] = 1; p0[4 * CS] = 0;
            if (GST) { double* p1 = M.stw; p1[0] = 0; p1[CS] = 0; p1[2 * CS] = 0; p1[3 * CS] = 1; p1[4 * CS] = 0; }
        }
        double p[2] = {0, 0};
        double s, c; ipc_sincos(th0, &s, &c);
        ts.pa.t = th0; ts.ca = c; ts.sa = s;
        auto rec_of = [&](int k) -> const double* {         // record of local edge k (this thread's step k - k0)
            return STG ? O.sm + 3 * k : (ODOM_DIRECT ? odom_rec<UNI>(O, k) : O.zs + (UNI ? 3 : 9) * (size_t)((k - k0) * NT + tid));
        };
        for (int kb = k0; kb < k1; kb += 4) {
            double zx[4], zy[4], zt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const double* zr = rec_of(kb + u < k1 ? kb + u : k1 - 1); zx[u] = zr[0]; zy[u] = zr[1]; zt[u] = zr[2]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u;
                if (k < k1) {
                    p[0] += c * zx[u] - s * zy[u]; p[1] += s * zx[u] + c * zy[u];
                    acc += zt[u];
                    const double thk = wrap_pi_hd(acc);
                    ipc_sincos(thk, &s, &c);
                    double* pq = SA::step(M, k + 1, k - k0, tid);
                    pq[2 * CS] = thk; pq[3 * CS] = c; pq[4 * CS] = s;
                }
            }
        }
        bsync<NT>();
        hd_block_excl_scan<NT, 2>(p, M.red2());
        double ax = p[0], ay = p[1];
        ts.pa.x = ax; ts.pa.y = ay;
        c = ts.ca; s = ts.sa;
        for (int kb = k0; kb < k1; kb += 4) {
            double zx[4], zy[4], cn[4], sn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u < k1 ? kb + u : k1 - 1;
                const double* zr = rec_of(k); zx[u] = zr[0]; zy[u] = zr[1];
                const double* pq = SA::step(M, k + 1, k - k0, tid); cn[u] = pq[3 * CS]; sn[u] = pq[4 * CS];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u;
                if (k < k1) {
                    ax += c * zx[u] - s * zy[u]; ay += s * zx[u] + c * zy[u];
                    double* pq = SA::step(M, k + 1, k - k0, tid);
                    pq[0] = ax; pq[CS] = ay; c = cn[u]; s = sn[u];
                }
            }
        }
#pragma unroll
        for (int m = 0; m < NPRE; ++m) ts.base[m] = 0;
    }

    int buf = 0, n_sweeps = 0, n_norm = 0, n_sd = 0, n_relin = 0, n_blend = 0;
#ifdef IPC_PHASE_CLOCKS
    long long ph_last = 0;
#ifdef __CUDA_ARCH__
    ph_last = ph_start;
#endif
    IPC_PH(0);
#endif
    SweepOut so;
    double cur_chi = 0, cur_max = 0, cand_chi = 0;
    double delta = 1e4;
    int iterations = 0, evals = 0, it = 0, tries = 0;
    double prev_hnorm = -1;      // norm of the last accepted step (speculation heuristic)
    bool have_norm = false, have_sd = false, need_rollback = false;
    bool last_bound = false;     // the previous decision was trust-region bound (steepest-descent / blend step)
    double hgnNorm = 0, linearGain = 0;
    double* sc = M.U()->sc;
    double &bb = sc[0], &bh = sc[1], &hh = sc[2], &bHb = sc[3], &alpha = sc[4], &hsdNorm = sc[5];
    double& gain_loops = sc[6];      // loop part of the GN predicted gain of the CURRENT linearisation (M.U()->sol)
    int purpose = P_INIT, mode = STEP_NONE;
    double c1 = 0, c2 = 0;
    for (;;) {
        if (need_rollback) { rollback<NT, GST>(M, ts); need_rollback = false; IPC_PH(6); }
        IPC_PH(7);
        sweep<NT, UNI, STG, GST>(M, O, mode, c1, c2, ts, so, buf, spec_v IPC_PH_PASS); ++n_sweeps;
        IPC_PH_COUNT(mode == STEP_GN ? 8 : (mode == STEP_BLEND ? 9 : 10), 1);
        if (mode == STEP_NONE && purpose != P_INIT) ++n_relin;
        if (mode == STEP_BLEND) ++n_blend;
        bool start_iter = false, after_reject = false, decide = false;
        if (purpose == P_INIT) {
            double n_c, n_m;
            eval_and_solve<NT>(M, buf, so.chi, 0, 1, true, n_c, n_m);
            IPC_PH(3);
            cur_chi = so.chi + n_c + n_m; cur_max = fmax(so.mx, fmax(n_c, n_m)); cand_chi = n_c;
            gain_loops = M.U()->sol.gain_loops;
            start_iter = true;
        } else if (purpose == P_RELIN_SPECFAIL) {
            decide = true;                                   // same try, the GN norm is known now
        } else if (purpose == P_RELIN_REJECT) {
            after_reject = true;
        } else {
            bool specfail = false;
            if (purpose == P_TRIAL_SPEC) { hgnNorm = sqrt(so.hh); have_norm = true; specfail = !(hgnNorm < delta); }
            if (specfail) {      // the GN step does not fit the trust region after all: undo, restore the linearisation
                need_rollback = true; mode = STEP_NONE; purpose = P_RELIN_SPECFAIL;
                continue;
            }
            const bool trial_gn = purpose != P_TRIAL_BLEND;
            const double hdlNorm = sqrt(so.hh);
            if (trial_gn) linearGain = so.gain + gain_loops;     // h_gn^T H h_gn (= b^T h_gn = chi2 - model for the exact GN step)
            ++evals;
            // loop edges at the trial state; thread 0 also solves the new linearisation when the step is going to be kept
            double n_c, n_m;
            eval_and_solve<NT>(M, buf, so.chi, cur_chi, linearGain, false, n_c, n_m);
            IPC_PH(3);
            const double newChi = so.chi + n_c + n_m;
            const double rawGain = linearGain;
            double lg = linearGain;
            if (fabs(lg) < 1e-12) lg = 1e-12;
            const double rho = (cur_chi - newChi) / lg;
            if (rho > 0.75) delta = fmax(delta, 3 * hdlNorm);
            else if (rho < 0.25) delta *= 0.5;
            if (rho > 0) {
                if (GST) { double* t_ = M.st; M.st = M.stw; M.stw = t_; }     // the trial state becomes the state
                cur_chi = newChi; cur_max = fmax(so.mx, fmax(n_c, n_m)); cand_chi = n_c;
                gain_loops = M.U()->sol.gain_loops;
                prev_hnorm = hdlNorm;
                ++iterations; ++it;
                start_iter = true;
            } else {
                need_rollback = true;
                prev_hnorm = -1;
                // a rejected Gauss-Newton step is retried verbatim while it still fits the trust region: every such retry
                // reproduces the same rho (<= 0), so only the halving of delta and the try counter advance
                if (trial_gn) while (tries < prm.max_tries && hgnNorm < delta) { ++tries; ++evals; delta *= 0.5; }
                if (prm.noise_eps > 0 && rawGain <= prm.noise_eps * cur_chi + 1e-300) tries = prm.max_tries;
                if (trial_gn && tries < prm.max_tries) {     // prefixes of the old state are needed again (gradient sweeps)
                    mode = STEP_NONE; purpose = P_RELIN_REJECT;
                    continue;
                }
                after_reject = true;
            }
        }
        if (after_reject) {
            if (tries < prm.max_tries) decide = true;
            else { ++iterations; break; }                    // Terminate: no good step in max_tries tries
            ++tries;
        }
        if (start_iter) {
            if (it >= max_iter) break;
            if (prm.early_accept && !want_info && cur_chi <= th) break;   // every edge chi2 <= sum <= th, and the sum only decreases
            have_norm = false; have_sd = false; tries = 1;
            decide = true;
        }
        if (decide) {
            if (!have_norm) {
                if (prm.speculate && prev_hnorm >= 0 && 4 * prev_hnorm < delta) {
                    mode = STEP_GN; purpose = P_TRIAL_SPEC; c1 = 0; c2 = 1;
                    continue;
                }
                if (need_rollback) { rollback<NT, GST>(M, ts); need_rollback = false; IPC_PH(6); }
                IPC_PH(7);
                if (prm.sd_fuse == 1 || (prm.sd_fuse == 2 && last_bound)) {
                    sd_fused<NT, UNI, STG, GST>(M, O, ts, bb, bh, hh, bHb); ++n_sweeps; ++n_sd;
                    IPC_PH(5); IPC_PH_COUNT(12, 1);
                    hgnNorm = sqrt(hh);
                    alpha = bb / bHb; hsdNorm = alpha * sqrt(bb); have_sd = true;
                } else {
                    hgnNorm = sqrt(gn_norm_sq<NT, UNI, STG, GST>(M, O, ts)); ++n_norm; ++n_sweeps;
                    IPC_PH(4); IPC_PH_COUNT(11, 1);
                }
                have_norm = true;
            }
            if (hgnNorm < delta) {
                mode = STEP_GN; purpose = P_TRIAL_GN; c1 = 0; c2 = 1; last_bound = false;
                continue;
            }
            last_bound = true;
            if (!have_sd) {
                if (need_rollback) { rollback<NT, GST>(M, ts); need_rollback = false; IPC_PH(6); }
                IPC_PH(7);
                sd_fused<NT, UNI, STG, GST>(M, O, ts, bb, bh, hh, bHb); ++n_sweeps;
                IPC_PH(5); IPC_PH_COUNT(12, 1);
                ++n_sd;
                alpha = bb / bHb;
                hsdNorm = alpha * sqrt(bb);
                have_sd = true;
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
            // H h_gn = b  =>  h^T H h = c1^2 bHb + 2 c1 c2 bb + c2^2 bh,  b^T h = c1 bb + c2 bh
            linearGain = -(c1 * c1 * bHb + 2 * c1 * c2 * bb + c2 * c2 * bh) + 2 * (c1 * bb + c2 * bh);
            mode = STEP_BLEND; purpose = P_TRIAL_BLEND;
            continue;
        }
        break;
    }
    res.verdict = (cur_max > th) ? 0 : 1;                      // every edge chi2 <= th (src/consensus_utils.cpp:17-19)
    res.max_chi2 = cur_max; res.cand_chi2 = cand_chi; res.sum_chi2 = cur_chi;
    res.iterations = iterations; res.evals = evals; res.window_len = L; res.n_loops = K; res.n_sweeps = n_sweeps;
    res.n_norm = n_norm; res.n_sd = n_sd; res.n_relin = n_relin; res.n_blend = n_blend;
    IPC_PH(7); IPC_PH_COUNT(13, 1); IPC_PH_COUNT(14, L); IPC_PH_COUNT(15, evals);
}

}  // namespace ipcb
