// common.cuh — shared device helpers for the IPC check kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ipc_b200.h"

namespace ipcb {

// ---- loop-candidate record in HBM (128 B SE2), information already moved into the
// ---- "relative-pose" frame of the edge (see DESIGN.md "Edge frames") ---------------------------
struct LoopRec2 {
    int from, to;
    double meas[3];       // x y theta
    double D[6];          // E^T Omega E upper triangle (00 01 02 11 12 22)
    double V[6];          // D^-1
};

struct BatchArgs {
    const double* odom;        // AoS odometry records: 3 doubles / edge (UNI kernels) or 9 doubles / edge
    int n_pad;
    double Du[6], Vu[6];       // uniform-information specialisation (kernels instantiated with UNI = true)
    const void* loops;         // LoopRec2 / LoopRec3
    const int* member;         // per check: member loop index or -1
    const int* cand;           // per check: candidate loop index
    const int* work;           // check ids handled by this launch (sorted by window length, longest first)
    const int* n_work;         // device counter: number of entries in `work`
    int* next;                 // device counter: next unclaimed entry of `work` (CTAs claim checks dynamically)
    double* scratch;           // per-CTA global scratch (pose backup, gradient, h_gn; MODE 1: the state arrays too)
    size_t scratch_stride;     // doubles per CTA
    int Lcap;                  // capacity (edges) of the per-vertex arrays, even
    double fast_th, slow_th;
    int fast_iter, slow_iter;
    double noise_eps;          // see CheckParams::noise_eps
    int max_tries;
    int speculate;             // apply the GN step before its norm is known when the trust region is far away
    int sd_fuse;               // see CheckParams::sd_fuse
    int early_accept;          // verdict-only batches: stop once sum chi2 <= th (the verdict can no longer change)
    unsigned char* verdict;    // per check
    ipc_check_info* info;      // per check (may be null)
};

}  // namespace ipcb
