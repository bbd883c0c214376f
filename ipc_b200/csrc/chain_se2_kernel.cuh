// chain_se2_kernel.cuh — CUDA entry of the SE(2) window check (see chain_se2.cuh for the algorithm).
// MODE 0: per-vertex state (pose + cos / sin of the heading: 5 doubles / vertex) in shared memory;
// MODE 1: state in the per-CTA global scratch (windows longer than shared memory holds);
// MODE 2 (uniform-information graphs): MODE 0 + the window's odometry records staged in shared memory by one bulk asynchronous
//         copy per check (cp.async.bulk + mbarrier, the 1-D TMA path): 64 B of shared memory per vertex instead of 40.
#pragma once
#include "chain_se2.cuh"

namespace ipcb {

// MINB = CTAs per SM the kernel is compiled for: registers per thread <= 65536 / (NT * MINB). The edge loop wants ~170
// registers; variants with fewer threads and no spills beat fuller SMs that spill (profiles/).
template <int NT, int MODE, bool UNI, int MINB>
__global__ void __launch_bounds__(NT, MINB) chain_check_se2(BatchArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int capv = A.Lcap + 2;
    double* scr = A.scratch + (size_t)blockIdx.x * A.scratch_stride;
    ChainMem M;
    M.small = sm; M.scr = scr; M.capv = capv; M.capg = scratch_slots<NT>(capv);
    M.st = (MODE != 1) ? sm + CHAIN_SMALL_DOUBLES : scr + (size_t)CHAIN_SCRATCH_ARRAYS * M.capg;
    M.stw = (MODE != 1) ? M.st : M.st + global_state_doubles(capv, NT);
    M.ring = (MODE == 1) ? sm + CHAIN_SMALL_DOUBLES : nullptr;
    StageMem stg{nullptr, nullptr, 0u};
    if (MODE == 2) {
        stg.buf = sm + CHAIN_SMALL_DOUBLES + (size_t)CHAIN_STATE_ARRAYS * capv;
        stg.mbar = reinterpret_cast<unsigned long long*>(stg.buf + stage_doubles(capv));
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(stg.mbar)), "r"(1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const LoopRec2* loops = static_cast<const LoopRec2*>(A.loops);
    CheckParams prm{A.fast_th, A.slow_th, A.fast_iter, A.slow_iter, A.noise_eps, A.max_tries, A.speculate, A.early_accept, A.sd_fuse};
    const int n_work = *A.n_work;
    __shared__ int s_wi;
    for (;;) {
        // dynamic claim: checks differ by 30x in cost (window length x Dogleg iterations), a static split leaves SMs idle
        if (threadIdx.x == 0) s_wi = atomicAdd(A.next, 1);
        if (NT <= 32) __syncwarp(); else __syncthreads();
        const int wi = s_wi;
        if (wi >= n_work) break;
        const int chk = A.work[wi];
        const int midx = A.member[chk];
        CheckResult r;
        run_check<NT, UNI, MODE == 2, MODE == 1>(M, A.odom, A.Du, A.Vu, loops + A.cand[chk], midx >= 0 ? loops + midx : nullptr, prm, A.info != nullptr, r, &stg);
        if (threadIdx.x == 0) {
            A.verdict[chk] = (unsigned char)r.verdict;
            if (A.info) {
                ipc_check_info o;
                o.max_chi2 = r.max_chi2; o.cand_chi2 = r.cand_chi2; o.sum_chi2 = r.sum_chi2;
                o.iterations = r.iterations; o.evals = r.evals; o.window_len = r.window_len; o.n_loops = r.n_loops;
                A.info[chk] = o;
            }
        }
    }
}

}  // namespace ipcb
