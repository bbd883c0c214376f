// launch_common.hpp — launch-shape tables and sizing helpers shared by the translation units of libipc_b200.so
// (ipc_capi.cu: C ABI + host logic; launch_se2.cu / launch_se3.cu: the chain-check kernel variants, built in parallel).
#pragma once
#include <algorithm>
#include <string>

#include "chain_se2.cuh"
#include "chain_se3.cuh"

namespace ipcb {
int fail(int code, const std::string& msg);
}
#define CUDA_TRY(x)                                                                                                   \
    do {                                                                                                              \
        cudaError_t _e = (x);                                                                                         \
        if (_e != cudaSuccess) return ipcb::fail(IPC_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e));      \
    } while (0)

namespace ipcb {

constexpr int NB = 6;   // launch buckets
struct Bucket { int cap; int nt; int mode; int minb; };   // minb: CTAs per SM the variant is compiled for (register budget)

inline size_t smem_bytes(int mode, int cap, int dim = 2, int nt = 0) {
    size_t capv = std::max(cap + 2, nt);
    size_t n = dim == 2 ? CHAIN_SMALL_DOUBLES + (mode != 1 ? (size_t)CHAIN_STATE_ARRAYS * capv : (size_t)ring_doubles(nt > 0 ? nt : 512)) + (mode == 2 ? (size_t)stage_doubles((int)capv) + 2 : 0)
                        : se3::CHAIN3_SMALL_DOUBLES + (mode == 0 ? (size_t)se3::CHAIN3_STATE * capv : 0);
    return n * sizeof(double);
}
inline size_t scratch_doubles_per_cta(int mode, int cap, int dim = 2, int nt = 0) {
    size_t capv = std::max(cap + 2, nt);
    // SE(2): the scratch is indexed by slot (vslot, chain_se2.cuh): up to capv + 2 nt + 2 records; nt = 0 sizes for the widest CTA
    if (mode == 1) nt = 512;                        // global-state kernels: sized for the widest CTA (32 .. 512 threads, launch_se2_variant)
    const size_t capg = capv + 2 * (size_t)(nt > 0 ? nt : 512) + 2;
    return dim == 2 ? (size_t)CHAIN_SCRATCH_ARRAYS * capg + (mode == 1 ? 2 * (size_t)global_state_doubles((int)capv, 512) : 0)
                    : (size_t)se3::CHAIN3_SCRATCH * capv + (mode == 1 ? 2 * (size_t)se3::global_state3_doubles((int)capv, 512) : 0);
}
// the instantiated (threads, CTAs per SM) variants of the SE(2) kernel, and the SE(3) variants by thread count
int launch_se2_variant(int nt, int minb, int mode, const BatchArgs& a, int grid, cudaStream_t st, bool uni);
int launch_se3_variant(int nt, int mode, const BatchArgs& a, int grid, cudaStream_t st);

}  // namespace ipcb
