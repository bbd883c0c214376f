// common.cuh — shared device helpers for the IPC check kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ipc_b200.h"

namespace ipcb {

// ---- loop-candidate record in HBM (80 B SE2 / 240 B SE3), information already moved into the
// ---- "relative-pose" frame of the edge (see DESIGN.md "Edge frames") ---------------------------
struct LoopRec2 {
    int from, to;
    double meas[3];       // x y theta
    double D[6];          // E^T Omega E upper triangle (00 01 02 11 12 22)
};

struct BatchArgs {
    const double* odom;        // SoA: NCOMP component arrays of length n_pad
    int n_pad;
    const void* loops;         // LoopRec2 / LoopRec3
    const int* member;         // per check: member loop index or -1
    const int* cand;           // per check: candidate loop index
    const int* work;           // check ids handled by this launch (sorted by window length, longest first)
    const int* n_work;         // device counter: number of entries in `work`
    double* scratch;           // MODE 2: per-CTA state arrays
    int Lcap;                  // capacity (edges) of the shared-memory arrays, even
    double fast_th, slow_th;
    int fast_iter, slow_iter;
    int noise_exit;
    int max_tries;
    unsigned char* verdict;    // per check
    ipc_check_info* info;      // per check (may be null)
};

// ---- g2o normalize_theta: [-pi, pi) ---------------------------------------------------------------
__device__ __forceinline__ double wrap_pi(double t) {
    const double pi = 3.14159265358979323846;
    if (t >= -pi && t < pi) return t;
    double m = floor(t / (2 * pi));
    t = t - m * 2 * pi;
    if (t >= pi) t -= 2 * pi;
    if (t < -pi) t += 2 * pi;
    return t;
}

template <int NT> __device__ __forceinline__ void bsync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}

// sum of M values over the block, result identical in every thread (xor butterfly + fixed-order cross-warp sum)
template <int NT, int M> __device__ __forceinline__ void block_sum(double (&v)[M], double* red) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[m] += __shfl_xor_sync(0xffffffffu, v[m], o);
    }
    if (NT == 32) return;
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
#pragma unroll
        for (int i = 0; i < NW; ++i) s += red[i * M + m];
        v[m] = s;
    }
}
template <int NT> __device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (NT == 32) return v;
    const int w = threadIdx.x >> 5, NW = NT / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) s = fmax(s, red[i]);
    return s;
}
// exclusive prefix over threads of M values (thread t receives sum over threads < t); v is replaced
template <int NT, int M> __device__ __forceinline__ void block_excl_scan(double (&v)[M], double* red) {
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
}

// ---- mbarrier + 1-D bulk copy (TMA engine, UBLKCP in SASS) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace ipcb
