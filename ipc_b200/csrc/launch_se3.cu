// launch_se3.cu — instantiations of the SE(3) chain-check kernel (own translation unit)
#include <atomic>

#include "launch_common.hpp"
#include "chain_se3_kernel.cuh"

namespace ipcb {

template <int NT, int MODE> int launch_se3(const BatchArgs& a, int grid, cudaStream_t st) {
    size_t sm = smem_bytes(MODE, a.Lcap, 3, NT);
    // the opt-in is per device (a process may hold handles on several GPUs): one flag per ordinal, set under the launch that needs it
    static std::atomic<bool> attr_done[64];
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev].load(std::memory_order_acquire)) {
        CUDA_TRY(cudaFuncSetAttribute(chain_check_se3<NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        if (dev >= 0 && dev < 64) attr_done[dev].store(true, std::memory_order_release);
    }
    chain_check_se3<NT, MODE><<<grid, NT, sm, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return IPC_OK;
}
int launch_se3_variant(int nt, int mode, const BatchArgs& a, int grid, cudaStream_t st) {
    if (mode == 1) {      // state in global memory (step tiles, second buffer)
        if (nt == 32) return launch_se3<32, 1>(a, grid, st);
        if (nt == 64) return launch_se3<64, 1>(a, grid, st);
        return launch_se3<256, 1>(a, grid, st);
    }
    if (nt == 32) return launch_se3<32, 0>(a, grid, st);
    if (nt == 64) return launch_se3<64, 0>(a, grid, st);
    if (nt == 128) return launch_se3<128, 0>(a, grid, st);
    if (nt == 256) return launch_se3<256, 0>(a, grid, st);
    return fail(IPC_ERR_ARG, "no SE(3) kernel variant for " + std::to_string(nt) + " threads");
}

}  // namespace ipcb
