// launch_se2.cu — instantiations of the SE(2) chain-check kernel (own translation unit: built in parallel with the rest)
#include <atomic>

#include "launch_common.hpp"
#include "chain_se2_kernel.cuh"

namespace ipcb {

template <int NT, int MODE, bool UNI, int MINB> int launch_se2u(const BatchArgs& a, int grid, cudaStream_t st) {
    size_t sm = smem_bytes(MODE, a.Lcap, 2, NT);
    // the opt-in is per device (a process may hold handles on several GPUs): one flag per ordinal, set under the launch that needs it
    static std::atomic<bool> attr_done[64];
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev].load(std::memory_order_acquire)) {
        CUDA_TRY(cudaFuncSetAttribute(chain_check_se2<NT, MODE, UNI, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        if (dev >= 0 && dev < 64) attr_done[dev].store(true, std::memory_order_release);
    }
    chain_check_se2<NT, MODE, UNI, MINB><<<grid, NT, sm, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return IPC_OK;
}
template <int NT, int MODE, int MINB> int launch_se2(const BatchArgs& a, int grid, cudaStream_t st, bool uni) {
    return uni ? launch_se2u<NT, MODE, true, MINB>(a, grid, st) : launch_se2u<NT, MODE, false, MINB>(a, grid, st);
}
// the instantiated (threads, CTAs per SM) variants
int launch_se2_variant(int nt, int minb, int mode, const BatchArgs& a, int grid, cudaStream_t st, bool uni) {
    if (mode == 1) {      // state in the global scratch
        if (nt == 32) return launch_se2<32, 1, 8>(a, grid, st, uni);
        if (nt == 64) return launch_se2<64, 1, 4>(a, grid, st, uni);
        return nt == 256 ? launch_se2<256, 1, 1>(a, grid, st, uni) : launch_se2<512, 1, 1>(a, grid, st, uni);
    }
    if (mode == 2) {      // staged odometry (cp.async.bulk): uniform-information kernels only
        if (!uni) return fail(IPC_ERR_ARG, "staged-odometry kernels need a uniform-information graph");
#define VS(NT_, MB_) if (nt == NT_ && minb == MB_) return launch_se2u<NT_, 2, true, MB_>(a, grid, st);
        VS(32, 16) VS(64, 8) VS(128, 3) VS(128, 2) VS(256, 1)
#undef VS
        return fail(IPC_ERR_ARG, "no staged kernel variant for " + std::to_string(nt) + " threads x " + std::to_string(minb) + " CTAs per SM");
    }
#define V(NT_, MB_) if (nt == NT_ && minb == MB_) return launch_se2<NT_, 0, MB_>(a, grid, st, uni);
    V(32, 16) V(32, 8) V(64, 8) V(64, 4) V(64, 2) V(128, 4) V(128, 3) V(128, 2) V(128, 1) V(192, 2) V(256, 2) V(256, 1) V(384, 1) V(512, 1)
#undef V
    return fail(IPC_ERR_ARG, "no kernel variant for " + std::to_string(nt) + " threads x " + std::to_string(minb) + " CTAs per SM");
}

}  // namespace ipcb

#ifdef IPC_PHASE_CLOCKS
// profiling builds only: cycles per phase summed over the CTAs' thread 0 (see IPC_PH in chain_se2.cuh)
extern "C" int ipc_debug_phase_clocks(unsigned long long* out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, ipcb::g_phase, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(ipcb::g_phase, z, sizeof z) != cudaSuccess) return -1; }
    return 0;
}
#endif
