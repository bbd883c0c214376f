// chain_se3_kernel.cuh — CUDA entry of the SE(3) window check (see chain_se3.cuh).
// MODE 0: per-vertex state (t, q: 7 doubles / vertex) in shared memory; MODE 1: state in the per-CTA global scratch.
#pragma once
#include "chain_se3.cuh"

namespace ipcb {

template <int NT, int MODE>
__global__ void __launch_bounds__(NT, 256 / NT > 0 ? 256 / NT : 1) chain_check_se3(BatchArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int capv = A.Lcap + 2 > NT ? A.Lcap + 2 : NT;      // the dead-reckoning scan stages NT poses in the state array
    double* scr = A.scratch + (size_t)blockIdx.x * A.scratch_stride;
    se3::ChainMem3 M;
    M.small = sm; M.scr = scr; M.capv = capv;
    M.st = (MODE == 0) ? sm + se3::CHAIN3_SMALL_DOUBLES : scr + (size_t)se3::CHAIN3_SCRATCH * capv;
    M.gst = (MODE == 1); M.nt = NT; M.S = 1; M.cs = (MODE == 1) ? NT : 1;
    M.stw = (MODE == 1) ? M.st + se3::global_state3_doubles(capv, NT) : M.st;
    const se3::LoopRec3* loops = static_cast<const se3::LoopRec3*>(A.loops);
    CheckParams prm{A.fast_th, A.slow_th, A.fast_iter, A.slow_iter, A.noise_eps, A.max_tries, A.speculate, A.early_accept, A.sd_fuse};
    const int n_work = *A.n_work;
    __shared__ int s_wi;
    for (;;) {
        if (threadIdx.x == 0) s_wi = atomicAdd(A.next, 1);
        if (NT <= 32) __syncwarp(); else __syncthreads();
        const int wi = s_wi;
        if (wi >= n_work) break;
        const int chk = A.work[wi];
        const int midx = A.member[chk];
        CheckResult r;
        se3::run_check3<NT>(M, A.odom, loops + A.cand[chk], midx >= 0 ? loops + midx : nullptr, prm, A.info != nullptr, r);
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
