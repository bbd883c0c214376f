// ipc_capi.cu — C ABI (include/ipc_b200.h) over the CUDA kernels. Host side of the handle:
// IPC<EDGE,VERTEX> state of /root/reference/include/ipc/consensus.hpp:23-32 kept as flat arrays,
// with the graph resident in HBM. No CPU fallback: every compute entry point launches kernels.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <vector>

#include "handle.hpp"
#include "matrix.cuh"

namespace ipcb {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
}  // namespace ipcb


// ------------------------------------------------------------------------------------------------
// planning kernels: bucket checks by window length (K3: interval overlap) and pack verdict bits
// ------------------------------------------------------------------------------------------------
namespace ipcb {

__global__ void plan_checks(const int* __restrict__ lfrom_to, int loop_stride_ints, int n_checks, const int* __restrict__ member,
                            const int* __restrict__ cand, const int* __restrict__ bucket_cap, int n_buckets, int* __restrict__ counts,
                            int* __restrict__ work, int work_stride, unsigned long long* __restrict__ stats) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long sl = 0, sk = 0;
    if (c < n_checks) {
        const int* lc = lfrom_to + (size_t)cand[c] * loop_stride_ints;
        int ca = min(lc[0], lc[1]), cb = max(lc[0], lc[1]);
        int lo = ca, hi = cb, K = 1;
        int m = member[c];
        if (m >= 0) {
            const int* lm = lfrom_to + (size_t)m * loop_stride_ints;
            int ma = min(lm[0], lm[1]), mb = max(lm[0], lm[1]);
            if (min(mb, cb) - max(ma, ca) > 0) { K = 2; lo = min(ca, ma); hi = max(cb, mb); }   // src/consensus.cpp:157-159
        }
        int L = hi - lo;
        int b = 0;
        while (b < n_buckets - 1 && L > bucket_cap[b]) ++b;
        int pos = atomicAdd(&counts[b], 1);
        work[(size_t)b * work_stride + pos] = c;
        sl = (unsigned long long)L; sk = (unsigned long long)K;
    }
    // warp-aggregate the statistics
    for (int o = 16; o > 0; o >>= 1) { sl += __shfl_xor_sync(0xffffffffu, sl, o); sk += __shfl_xor_sync(0xffffffffu, sk, o); }
    if ((threadIdx.x & 31) == 0 && (sl | sk)) { atomicAdd(&stats[0], sl); atomicAdd(&stats[1], sk); }
}

__global__ void pack_bits(const unsigned char* __restrict__ verdict, int n, uint32_t* __restrict__ bits) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v = (c < n) ? verdict[c] : 0;
    unsigned w = __ballot_sync(0xffffffffu, v != 0);
    if ((threadIdx.x & 31) == 0 && c < n) bits[c >> 5] = w;
}

// sharded batches: rank r owns the checks r, r + world, ... of a list
__global__ void shard_take(const int* __restrict__ member, const int* __restrict__ cand, int rank, int world, int n_local, int* __restrict__ lm,
                           int* __restrict__ lc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_local) { const size_t c = (size_t)rank + (size_t)i * world; lm[i] = member[c]; lc[i] = cand[c]; }
}
// gathered words [world][wpr] -> verdict byte per check of the whole list
__global__ void shard_spread(const uint32_t* __restrict__ gathered, size_t wpr, int world, int n_checks, unsigned char* __restrict__ verdict) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_checks) { const int r = c % world, i = c / world; verdict[c] = (gathered[(size_t)r * wpr + (i >> 5)] >> (i & 31)) & 1u; }
}

}  // namespace ipcb

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
namespace {

// Launch tables: window-length caps (edges), threads per CTA, memory mode (see chain_se2_kernel.cuh), CTAs per SM the variant is
// compiled for. The edge loop needs ~245 registers, so the variants that matter hold 256 threads per SM (8 x 32, 4 x 64, 2 x 128,
// 1 x 256); the 128-register variants (16 x 32, 8 x 64) only serve windows too short to matter. MODE 0 keeps 5 doubles per vertex in
// shared memory, MODE 1 streams the state from global memory (step tiles + cp.async ring + second buffer), MODE 2 adds the staged odometry.
// CTA-per-check table (general-information graphs; option cta_per_check):
const Bucket kBuckets2[NB] = {{96, 32, 0, 16}, {320, 64, 0, 8}, {1300, 128, 0, 3}, {2600, 128, 0, 2}, {5400, 256, 0, 1}, {1 << 30, 256, 1, 1}};

// Uniform-information SE(2) graphs (M3500, City10000, the 50 k config): ONE WARP PER CHECK, eight independent checks per SM. No block
// barriers, no idle warps while one lane solves the force system; windows up to 540 edges keep their state in shared memory, longer
// ones stream it from L2 / HBM in coalesced step tiles through a cp.async ring (TileFeed) and write trial states to a second buffer
// (no backup, no rollback pass). Measured against the CTA-per-check table on the M3500 sample (profiles/r02_ab_*): 1.73x for
// 320 < L <= 540, +16 % / +4 % / +33 % for 1300-2000 / 2000-2600 / 2600-3499, +24 % on the whole list; on the City10000-shaped matrix (windows up to 9997 edges)
// one warp per check up to 10000 edges gives +7.7 % over handing over at 5400 (profiles/r02_ab_log.txt). Windows beyond 10000 edges keep
// 256 threads per check (one check per SM) on the same streamed layout (50 k-pose config: no difference between the two within noise). The three middle rows share one shape: the split only evens
// out the tail of the dynamic work claim (longest windows first within a launch).
const Bucket kBuckets2U[NB] = {{96, 32, 0, 16}, {540, 32, 0, 8}, {1500, 32, 1, 8}, {2600, 32, 1, 8}, {10000, 32, 1, 8}, {1 << 30, 256, 1, 1}};

// MODE 2 (option stage_odom = 1, uniform-information graphs): + 24 B / vertex for the odometry window staged by cp.async.bulk: the
// same CTAs per SM hold shorter windows (measured A/B in profiles/)
const Bucket kBuckets2S[NB] = {{96, 32, 2, 16}, {320, 64, 2, 8}, {1100, 128, 2, 3}, {1700, 128, 2, 2}, {3500, 256, 2, 1}, {1 << 30, 256, 1, 1}};

// SE(3): 7 doubles of state per vertex, 256 resident threads per SM (the 27 running prefix values need the registers)
const Bucket kBuckets3[NB] = {{96, 32, 0, 8}, {320, 64, 0, 4}, {800, 128, 0, 2}, {3700, 256, 0, 1}, {3701, 256, 0, 1}, {1 << 30, 256, 1, 1}};
const Bucket* buckets_of(int dim, bool uni = false) { return dim == 2 ? (uni ? kBuckets2U : kBuckets2) : kBuckets3; }


}  // namespace

namespace { int size_scratch(ipc_handle* h); }
extern "C" { namespace { int stream_solver_setup(ipc_handle* h); void slot_free(ClSlot& s); } }


namespace {

// per-CTA scratch: every bucket has its own region (grid * stride doubles), because the bucket launches of one batch run side by
// side on their own streams; also uploads the bucket caps for plan_checks
int size_scratch(ipc_handle* h) {
    size_t need = 0;
    const Bucket* kB = h->buckets;
    for (int b = 0; b < NB; ++b) { h->scratch_off[b] = 0; h->scratch_len[b] = 0; }
    for (int b = 0; b < NB; ++b) {
        int lo_cap = b == 0 ? 0 : kB[b - 1].cap;
        if (lo_cap >= h->n - 1) break;
        int Lcap = (std::min(kB[b].cap, h->n - 1) + 1) & ~1;
        size_t sm = smem_bytes(kB[b].mode, Lcap, h->dim, kB[b].nt);
        if (kB[b].mode == 0 && sm > 226 * 1024) return fail(IPC_ERR_ARG, "bucket " + std::to_string(b) + " does not fit shared memory");
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(32, (227 * 1024) / (sm + 1024)));
        per_sm = std::min(per_sm, kB[b].minb);
        h->scratch_off[b] = need;
        h->scratch_len[b] = ((size_t)h->n_sm * per_sm * scratch_doubles_per_cta(kB[b].mode, Lcap, h->dim, kB[b].nt) + 31) & ~(size_t)31;   // 256-byte aligned regions
        need += h->scratch_len[b];
    }
    if (need > h->scratch_doubles) {
        cudaFree(h->d_scratch); h->d_scratch = nullptr; h->scratch_doubles = 0;
        CUDA_TRY(cudaMalloc(&h->d_scratch, need * sizeof(double)));
        h->scratch_doubles = need;
    }
    int caps[NB];
    for (int b = 0; b < NB; ++b) caps[b] = kB[b].cap;
    CUDA_TRY(cudaMemcpy(h->d_bucket_cap, caps, sizeof(caps), cudaMemcpyHostToDevice));
    return IPC_OK;
}

int ensure_batch_buffers(ipc_handle* h, int n_checks) {
    if (n_checks <= h->cap_checks) return IPC_OK;
    int cap = std::max(n_checks, 1024);
    cudaFree(h->d_member); cudaFree(h->d_cand); cudaFree(h->d_work); cudaFree(h->d_verdict); cudaFree(h->d_bits); cudaFree(h->d_info);
    h->d_member = h->d_cand = h->d_work = nullptr; h->d_verdict = nullptr; h->d_bits = nullptr; h->d_info = nullptr;
    h->cap_checks = 0;
    CUDA_TRY(cudaMalloc(&h->d_member, sizeof(int) * cap));
    CUDA_TRY(cudaMalloc(&h->d_cand, sizeof(int) * cap));
    CUDA_TRY(cudaMalloc(&h->d_work, sizeof(int) * (size_t)cap * NB));
    CUDA_TRY(cudaMalloc(&h->d_verdict, cap));
    CUDA_TRY(cudaMalloc(&h->d_bits, sizeof(uint32_t) * ((cap + 31) / 32)));
    CUDA_TRY(cudaMalloc(&h->d_info, sizeof(ipc_check_info) * cap));
    h->cap_checks = cap;
    return IPC_OK;
}

// enqueue plan + bucket launches + pack on `st`; all pointers device; work buffer sized for cap >= n_checks
int enqueue_batch(ipc_handle* h, int n_checks, const int* member_dev, const int* cand_dev, int* work_dev, int work_stride,
                  unsigned char* verdict_dev, uint32_t* bits_dev, ipc_check_info* info_dev, cudaStream_t st) {
    const Bucket* kB = h->buckets;
    CUDA_TRY(cudaMemsetAsync(h->d_counts, 0, sizeof(int) * 2 * NB, st));
    CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, sizeof(unsigned long long) * 2, st));
    const int loop_stride_ints = (int)((h->dim == 2 ? sizeof(LoopRec2) : sizeof(se3::LoopRec3)) / sizeof(int));
    plan_checks<<<(n_checks + 255) / 256, 256, 0, st>>>(reinterpret_cast<const int*>(h->d_loops), loop_stride_ints, n_checks, member_dev, cand_dev,
                                                        h->d_bucket_cap, NB, h->d_counts, work_dev, work_stride, h->d_stats);
    CUDA_TRY(cudaGetLastError());
    int launches = 1;
    CUDA_TRY(cudaEventRecord(h->ev_k0, st));
    // The buckets are independent (own work list, own counters, own scratch region, disjoint verdict slots): fork them onto one
    // stream each behind the plan, longest windows first, and join before the pack — the tail of one launch (a few long checks still
    // running) and an underfilled launch (few checks of one length class) overlap with the other buckets instead of idling SMs.
    const bool overlap = h->overlap_buckets != 0 && h->fork_ev != nullptr;
    if (overlap) CUDA_TRY(cudaEventRecord(h->fork_ev, st));
    int n_b = 0;
    while (n_b < NB && (n_b == 0 ? 0 : kB[n_b - 1].cap) < h->n - 1) ++n_b;        // buckets some window of this graph can fall into
    for (int q = 0; q < n_b; ++q) {
        const int b = overlap ? n_b - 1 - q : q;
        const Bucket& bk = kB[b];
        cudaStream_t bs = overlap ? h->bucket_stream[b] : st;
        if (overlap) CUDA_TRY(cudaStreamWaitEvent(bs, h->fork_ev, 0));
        BatchArgs a{};
        for (int c = 0; c < 6; ++c) { a.Du[c] = h->hs.Du[c]; a.Vu[c] = h->hs.Vu[c]; }
        const bool uni = h->dim == 2 && h->hs.uniform_iso && h->use_uniform && h->d_odom3;
        a.odom = h->dim == 3 ? h->d_odom49 : (uni ? h->d_odom3 : h->d_odom9); a.n_pad = h->n_pad; a.loops = h->d_loops; a.member = member_dev; a.cand = cand_dev;
        a.work = work_dev + (size_t)b * work_stride; a.n_work = h->d_counts + b; a.next = h->d_counts + NB + b;
        a.Lcap = (std::min(bk.cap, h->n - 1) + 1) & ~1;
        a.fast_th = h->cfg.fast_reject_th; a.slow_th = h->cfg.slow_reject_th;
        a.fast_iter = h->cfg.fast_reject_iter_base; a.slow_iter = h->cfg.slow_reject_iter_base;
        a.noise_eps = h->noise_eps; a.max_tries = h->max_tries; a.speculate = h->speculate; a.early_accept = h->early_accept; a.sd_fuse = h->sd_fuse;
        a.verdict = verdict_dev; a.info = info_dev; a.scratch = h->d_scratch + h->scratch_off[b];
        a.scratch_stride = scratch_doubles_per_cta(bk.mode, a.Lcap, h->dim, bk.nt);
        size_t sm = smem_bytes(bk.mode, a.Lcap, h->dim, bk.nt);
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(32, (227 * 1024) / (sm + 1024)));
        per_sm = std::min(per_sm, bk.minb);
        int grid = std::min(n_checks, h->n_sm * per_sm);
        grid = (int)std::min<size_t>(grid, h->scratch_len[b] / a.scratch_stride);
        if (grid < 1) return fail(IPC_ERR_CUDA, "scratch region of bucket " + std::to_string(b) + " is smaller than one CTA's stride");
        int rc = IPC_OK;
        if (h->dim == 3) {
            rc = launch_se3_variant(bk.nt, bk.mode, a, grid, bs);
        } else rc = launch_se2_variant(bk.nt, bk.minb, bk.mode, a, grid, bs, uni);
        if (rc != IPC_OK) return rc;
        if (overlap) {
            CUDA_TRY(cudaEventRecord(h->bucket_ev[b], bs));
            CUDA_TRY(cudaStreamWaitEvent(st, h->bucket_ev[b], 0));
        }
        ++launches;
    }
    CUDA_TRY(cudaEventRecord(h->ev_k1, st));
    h->ev_valid = true;
    if (bits_dev) {
        pack_bits<<<(n_checks + 255) / 256, 256, 0, st>>>(verdict_dev, n_checks, bits_dev);
        CUDA_TRY(cudaGetLastError());
        ++launches;
    }
    h->last_launches = launches;
    return IPC_OK;
}

}  // namespace

extern "C" {

const char* ipc_last_error(void) { return g_err.c_str(); }

int ipc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int ipc_create(int dim, int n_poses, const double* odom_meas, const double* odom_info, const ipc_config* cfg, int device, ipc_handle** out) {
    if (!out) return fail(IPC_ERR_ARG, "out is null");
    *out = nullptr;
    if (dim != 2 && dim != 3) return fail(IPC_ERR_ARG, "dim must be 2 or 3");
    if (n_poses < 2 || !odom_meas || !odom_info || !cfg) return fail(IPC_ERR_ARG, "need n_poses >= 2 and non-null odometry / cfg");
    int ndev = ipc_device_count();
    if (ndev <= 0) return fail(IPC_ERR_CUDA, "no CUDA device: this library has no CPU path");
    if (device < 0 || device >= ndev) return fail(IPC_ERR_ARG, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    ipc_handle* h = new ipc_handle();
    struct Guard { ipc_handle* h; ~Guard() { if (h) ipc_destroy(h); } } guard{h};     // any early return below frees the handle and its buffers
    h->dim = dim; h->d = dim == 2 ? 3 : 6; h->mw = dim == 2 ? 3 : 7;
    h->n = n_poses; h->n_pad = (n_poses + 3) & ~1; h->device = device; h->cfg = *cfg;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    h->n_sm = prop.multiProcessorCount;
    std::string err;
    if (!h->hs.init(dim, n_poses, odom_meas, odom_info, cfg->s_factor, err)) return fail(IPC_ERR_ARG, err);
    // device odometry records
    std::vector<double> rec;
    if (dim == 3) {
        if (!h->hs.build_odom_aos3(h->n_pad, rec)) { return fail(IPC_ERR_ARG, "odometry edge with a zero quaternion or a singular information matrix"); }
        CUDA_TRY(cudaMalloc(&h->d_odom49, rec.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(h->d_odom49, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
    } else {
        h->hs.build_odom_aos(false, h->n_pad, rec);
        CUDA_TRY(cudaMalloc(&h->d_odom9, rec.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(h->d_odom9, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (dim == 2 && h->hs.uniform_iso) {
        h->hs.build_odom_aos(true, h->n_pad, rec);
        CUDA_TRY(cudaMalloc(&h->d_odom3, rec.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(h->d_odom3, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMalloc(&h->d_counts, sizeof(int) * 2 * NB));
    CUDA_TRY(cudaMalloc(&h->d_bucket_cap, sizeof(int) * NB));
    CUDA_TRY(cudaMalloc(&h->d_stats, sizeof(unsigned long long) * 2));
    for (int b = 0; b < NB; ++b) h->buckets[b] = buckets_of(dim, dim == 2 && h->hs.uniform_iso && h->use_uniform && h->d_odom3)[b];
    { int rc2 = size_scratch(h); if (rc2 != IPC_OK) return rc2; }
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {   // IPC::IPC: vertex 0 at the origin, everything else dead-reckoned (propagateGuess, src/consensus_utils.cpp:98-116)
        CUDA_TRY(cudaMalloc(&h->d_pose, sizeof(double) * (dim == 2 ? 5 : 7) * (size_t)n_poses));
        if (dim == 2) {
            cl_set_origin<<<1, 32, 0, h->stream>>>(h->d_pose);
            cl_dead_reckon<<<1, CL_NT, 0, h->stream>>>(h->d_odom9, 0, n_poses, h->d_pose);
        } else {
            CUDA_TRY(cudaMalloc(&h->cl_stage, sizeof(double) * 7 * CL_NT));
            cl3_set_origin<<<1, 32, 0, h->stream>>>(h->d_pose);
            cl3_dead_reckon<<<1, CL_NT, 0, h->stream>>>(h->d_odom49, 0, n_poses, h->d_pose, h->cl_stage);
        }
        CUDA_TRY(cudaGetLastError());
        int rcs = stream_solver_setup(h);
        if (rcs != IPC_OK) return rcs;
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    CUDA_TRY(cudaEventCreate(&h->ev_k0));
    CUDA_TRY(cudaEventCreate(&h->ev_k1));
    for (int b = 0; b < NB; ++b) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->bucket_stream[b], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->bucket_ev[b], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
    guard.h = nullptr;
    *out = h;
    return IPC_OK;
}

void ipc_destroy(ipc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_odom9); cudaFree(h->d_odom3); cudaFree(h->d_odom49); cudaFree(h->d_loops); cudaFree(h->d_member); cudaFree(h->d_cand); cudaFree(h->d_work); cudaFree(h->d_counts);
    cudaFree(h->d_bucket_cap); cudaFree(h->d_verdict); cudaFree(h->d_bits); cudaFree(h->d_info); cudaFree(h->d_stats); cudaFree(h->d_scratch); cudaFree(h->d_gather);
    cudaFree(h->d_pose); cudaFree(h->d_odom9_raw); cudaFree(h->cl_res);
    cudaFree(h->cl_bar); cudaFree(h->cl_args); cudaFree(h->cl_prof); cudaFree(h->cl_out); cudaFree(h->cl_stage); cudaFree(h->d_odom49_raw);
    for (ClSlot& sl : h->slots) slot_free(sl);
    delete h->comm;
    for (int s = 0; s < CL_MAX_SLOTS; ++s) { if (h->slot_stream[s]) cudaStreamDestroy(h->slot_stream[s]); if (h->slot_ev[s]) cudaEventDestroy(h->slot_ev[s]); }
    if (h->commit_ev) cudaEventDestroy(h->commit_ev);
    if (h->h_abort) cudaFreeHost(h->h_abort);
    if (h->cl_hout) cudaFreeHost(h->cl_hout);
    if (h->cl_hargs) cudaFreeHost(h->cl_hargs);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int b = 0; b < NB; ++b) { if (h->bucket_stream[b]) cudaStreamDestroy(h->bucket_stream[b]); if (h->bucket_ev[b]) cudaEventDestroy(h->bucket_ev[b]); }
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    if (h->ev_k0) cudaEventDestroy(h->ev_k0);
    if (h->ev_k1) cudaEventDestroy(h->ev_k1);
    delete h;
}

int ipc_set_option(ipc_handle* h, const char* name, double value) {
    if (!h || !name) return fail(IPC_ERR_ARG, "null argument");
    if (!strcmp(name, "noise_exit")) { h->noise_eps = value == 1.0 ? 1e-13 : value; return IPC_OK; }   // 0 = off, 1 = default eps, else eps
    if (!strcmp(name, "sd_fuse")) { if (value < 0 || value > 2) return fail(IPC_ERR_ARG, "sd_fuse: 0, 1 or 2"); h->sd_fuse = (int)value; return IPC_OK; }
    if (!strncmp(name, "bucket", 6) && name[6] >= '0' && name[6] < '0' + NB && name[7] == '_') {
        const int b = name[6] - '0';
        if (!strcmp(name + 8, "cap")) h->buckets[b].cap = (int)value;
        else if (!strcmp(name + 8, "nt")) h->buckets[b].nt = (int)value;        // (nt, minb) must name an instantiated variant: checked at launch
        else if (!strcmp(name + 8, "minb")) h->buckets[b].minb = (int)value;
        else if (!strcmp(name + 8, "mode")) h->buckets[b].mode = (int)value;      // 0 state in shared memory, 1 in the global scratch, 2 staged odometry
        else return fail(IPC_ERR_ARG, std::string("unknown option ") + name);
        for (int q = 1; q < NB - 1; ++q) if (h->buckets[q].cap < h->buckets[q - 1].cap) return fail(IPC_ERR_ARG, "bucket caps must be non-decreasing");
        CUDA_TRY(cudaSetDevice(h->device));
        return size_scratch(h);
    }
    if (!strcmp(name, "stage_odom")) {     // 1: odometry window staged in shared memory by one bulk copy per check (MODE 2 kernels)
        if (value != 0 && !(h->dim == 2 && h->hs.uniform_iso && h->d_odom3)) return fail(IPC_ERR_UNSUPPORTED, "stage_odom needs an SE(2) graph with uniform isotropic odometry information");
        for (int b = 0; b < NB; ++b) h->buckets[b] = (value != 0 ? kBuckets2S : buckets_of(2, h->use_uniform))[b];
        CUDA_TRY(cudaSetDevice(h->device));
        return size_scratch(h);
    }
    if (!strcmp(name, "use_uniform")) {      // 0: general-information kernels (72-byte records) and their CTA-per-check launch table
        h->use_uniform = value != 0;
        if (h->dim == 2) {
            for (int b = 0; b < NB; ++b) h->buckets[b] = buckets_of(2, h->hs.uniform_iso && h->use_uniform && h->d_odom3)[b];
            CUDA_TRY(cudaSetDevice(h->device));
            return size_scratch(h);
        }
        return IPC_OK;
    }
    if (!strcmp(name, "cta_per_check")) {    // 1: the CTA-per-check launch table (state in shared memory) on a uniform-information graph too (A/B runs)
        if (h->dim != 2) return fail(IPC_ERR_UNSUPPORTED, "cta_per_check: SE(2) only");
        for (int b = 0; b < NB; ++b) h->buckets[b] = buckets_of(2, value == 0 && h->hs.uniform_iso && h->use_uniform && h->d_odom3)[b];
        CUDA_TRY(cudaSetDevice(h->device));
        return size_scratch(h);
    }
    if (!strcmp(name, "overlap_buckets")) { h->overlap_buckets = value != 0; return IPC_OK; }   // 0: the bucket launches of a batch one after the other on the caller's stream
    if (!strcmp(name, "speculate")) { h->speculate = value != 0; return IPC_OK; }
    if (!strcmp(name, "early_accept")) { h->early_accept = value != 0; return IPC_OK; }
    if (!strcmp(name, "max_tries")) { h->max_tries = (int)value; return IPC_OK; }
    if (!strcmp(name, "stream_depth")) { if (value < 1 || value > CL_MAX_SLOTS) return fail(IPC_ERR_ARG, "stream_depth: 1 .. 8"); h->stream_depth = (int)value; return IPC_OK; }
    return fail(IPC_ERR_ARG, std::string("unknown option ") + name);
}

int ipc_set_candidates(ipc_handle* h, int n_loops, const int* from, const int* to, const double* meas, const double* info) {
    if (!h || n_loops < 0 || (n_loops && (!from || !to || !meas || !info))) return fail(IPC_ERR_ARG, "bad candidate table");
    CUDA_TRY(cudaSetDevice(h->device));
    for (int i = 0; i < n_loops; ++i) {
        if (from[i] < 0 || to[i] < 0 || from[i] >= h->n || to[i] >= h->n || from[i] == to[i])
            return fail(IPC_ERR_ARG, "candidate " + std::to_string(i) + " has invalid vertex ids");
    }
    cudaFree(h->d_loops); h->d_loops = nullptr; h->n_loops = 0;
    h->h_lfrom.assign(from, from + n_loops); h->h_lto.assign(to, to + n_loops);
    h->h_lmeas.assign(meas, meas + (size_t)n_loops * h->mw); h->h_linfo.assign(info, info + (size_t)n_loops * h->d * h->d);
    if (h->dim == 2) {
        std::vector<LoopRec2> recs(n_loops);
        for (int i = 0; i < n_loops; ++i) {
            recs[i].from = from[i]; recs[i].to = to[i];
            HostState::se2_edge_record(meas + 3 * i, info + 9 * i, 1.0, recs[i].meas, recs[i].D); HostState::inv_sym3_host(recs[i].D, recs[i].V);
        }
        if (n_loops) {
            CUDA_TRY(cudaMalloc(&h->d_loops, sizeof(LoopRec2) * n_loops));
            CUDA_TRY(cudaMemcpy(h->d_loops, recs.data(), sizeof(LoopRec2) * n_loops, cudaMemcpyHostToDevice));
        }
    } else {
        std::vector<se3::LoopRec3> recs(n_loops);
        for (int i = 0; i < n_loops; ++i) {
            double r[49];
            if (!HostState::se3_edge_record(meas + 7 * i, info + 36 * i, 1.0, r)) return fail(IPC_ERR_ARG, "candidate " + std::to_string(i) + " has a zero quaternion or a singular information matrix");
            recs[i].from = from[i]; recs[i].to = to[i];
            for (int q = 0; q < 7; ++q) recs[i].zinv[q] = r[q];
            for (int q = 0; q < 21; ++q) { recs[i].Om[q] = r[7 + q]; recs[i].V[q] = r[28 + q]; }
        }
        if (n_loops) {
            CUDA_TRY(cudaMalloc(&h->d_loops, sizeof(se3::LoopRec3) * n_loops));
            CUDA_TRY(cudaMemcpy(h->d_loops, recs.data(), sizeof(se3::LoopRec3) * n_loops, cudaMemcpyHostToDevice));
        }
    }
    h->n_loops = n_loops;
    return IPC_OK;
}

int ipc_check_batch_dev(ipc_handle* h, int n_checks, const int* member_dev, const int* cand_dev, uint32_t* out_bits_dev,
                        ipc_check_info* out_info_dev, void* stream) {
    if (!h || n_checks < 0) return fail(IPC_ERR_ARG, "bad arguments");
    if (!h->d_loops) return fail(IPC_ERR_STATE, "no candidate table: call ipc_set_candidates first");
    if (n_checks == 0) return IPC_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = ensure_batch_buffers(h, n_checks);
    if (rc != IPC_OK) return rc;
    return enqueue_batch(h, n_checks, member_dev, cand_dev, h->d_work, h->cap_checks, h->d_verdict, out_bits_dev, out_info_dev,
                         static_cast<cudaStream_t>(stream));
}

int ipc_check_batch(ipc_handle* h, int n_checks, const int* member, const int* cand, uint32_t* out_bits, ipc_check_info* out_info) {
    if (!h || n_checks < 0 || (n_checks && (!member || !cand || !out_bits))) return fail(IPC_ERR_ARG, "bad arguments");
    if (!h->d_loops) return fail(IPC_ERR_STATE, "no candidate table: call ipc_set_candidates first");
    if (n_checks == 0) return IPC_OK;
    for (int i = 0; i < n_checks; ++i)
        if (cand[i] < 0 || cand[i] >= h->n_loops || member[i] >= h->n_loops) return fail(IPC_ERR_ARG, "check " + std::to_string(i) + " indexes outside the candidate table");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = ensure_batch_buffers(h, n_checks);
    if (rc != IPC_OK) return rc;
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->d_member, member, sizeof(int) * n_checks, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->d_cand, cand, sizeof(int) * n_checks, cudaMemcpyHostToDevice, st));
    rc = enqueue_batch(h, n_checks, h->d_member, h->d_cand, h->d_work, h->cap_checks, h->d_verdict, h->d_bits, out_info ? h->d_info : nullptr, st);
    if (rc != IPC_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_bits, h->d_bits, sizeof(uint32_t) * ((n_checks + 31) / 32), cudaMemcpyDeviceToHost, st));
    if (out_info) CUDA_TRY(cudaMemcpyAsync(out_info, h->d_info, sizeof(ipc_check_info) * n_checks, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return IPC_OK;
}

int ipc_last_batch_stats(ipc_handle* h, int64_t* sum_L, int64_t* sum_K, int* n_launches) {
    if (!h) return fail(IPC_ERR_ARG, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long s[2] = {0, 0};
    CUDA_TRY(cudaMemcpy(s, h->d_stats, sizeof(s), cudaMemcpyDeviceToHost));
    if (sum_L) *sum_L = (int64_t)s[0];
    if (sum_K) *sum_K = (int64_t)s[1];
    if (n_launches) *n_launches = h->last_launches;
    return IPC_OK;
}

int ipc_last_kernel_ms(ipc_handle* h, float* ms) {
    if (!h || !ms) return fail(IPC_ERR_ARG, "null argument");
    if (!h->ev_valid) return fail(IPC_ERR_STATE, "no batch has been enqueued");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->ev_k1));
    CUDA_TRY(cudaEventElapsedTime(ms, h->ev_k0, h->ev_k1));
    return IPC_OK;
}

// ---- consensus-set API: integer logic on the host mirror (src/consensus.cpp:77-171) ---------------
int ipc_remove_edge(ipc_handle* h, int from, int to, int* removed) {
    if (!h) return fail(IPC_ERR_ARG, "null handle");
    int r = h->hs.remove_edge(from, to) ? 1 : 0;
    if (removed) *removed = r;
    return IPC_OK;
}
int ipc_add_edge(ipc_handle* h, int from, int to, const double* meas, const double* info) {
    if (!h || !meas || !info) return fail(IPC_ERR_ARG, "null argument");
    if (from < 0 || to < 0 || from >= h->n || to >= h->n || from == to) return fail(IPC_ERR_ARG, "invalid vertex ids");
    h->hs.add_edge(from, to, meas, info);
    return IPC_OK;
}
int ipc_consensus_size(ipc_handle* h, int* n) {
    if (!h || !n) return fail(IPC_ERR_ARG, "null argument");
    *n = (int)h->hs.cns.size();
    return IPC_OK;
}
int ipc_get_consensus(ipc_handle* h, int* from_to, int capacity) {
    if (!h || !from_to) return fail(IPC_ERR_ARG, "null argument");
    if (capacity < (int)h->hs.cns.size()) return fail(IPC_ERR_ARG, "capacity too small");
    for (size_t i = 0; i < h->hs.cns.size(); ++i) { from_to[2 * i] = h->hs.cns[i].from; from_to[2 * i + 1] = h->hs.cns[i].to; }
    return IPC_OK;
}

namespace {

// the persistent stream-solver kernels (stream_solver.cuh): opt in to the shared memory they may need, size the cooperative grid
int stream_solver_setup(ipc_handle* h) {
    CUDA_TRY(cudaFuncSetAttribute(stream_check_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(stream_check_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    int coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) return fail(IPC_ERR_UNSUPPORTED, "device does not support cooperative launches");
    h->cl_grid = h->n_sm;          // one CTA of CL_NT threads per SM
    h->slots.resize(CL_MAX_SLOTS);
    CUDA_TRY(cudaMalloc(&h->cl_bar, sizeof(unsigned) * 4 * CL_MAX_SLOTS));
    CUDA_TRY(cudaMalloc(&h->cl_prof, sizeof(unsigned long long) * 16));
    CUDA_TRY(cudaMemset(h->cl_prof, 0, sizeof(unsigned long long) * 16));
    CUDA_TRY(cudaMalloc(&h->cl_out, sizeof(double) * 16 * CL_MAX_SLOTS));
    CUDA_TRY(cudaMallocHost(&h->cl_hout, sizeof(double) * 16 * CL_MAX_SLOTS));
    CUDA_TRY(cudaMalloc(&h->cl_args, sizeof(StreamArgs) * CL_MAX_SLOTS));
    CUDA_TRY(cudaMallocHost(&h->cl_hargs, sizeof(StreamArgs) * CL_MAX_SLOTS));
    CUDA_TRY(cudaMalloc(&h->cl_res, sizeof(double) * CL_NRES * CL_MAX_SLOTS));
    return IPC_OK;
}

void slot_free(ClSlot& s) {
    for (int q = 0; q < 2; ++q) { cudaFree(s.B[q].W); cudaFree(s.B[q].T); cudaFree(s.B[q].P); cudaFree(s.B[q].chi_e); cudaFree(s.B[q].lt); s.B[q] = ClBuffers{}; }
    cudaFree(s.G); cudaFree(s.H); cudaFree(s.S); cudaFree(s.z); cudaFree(s.lg); cudaFree(s.ev_ptr); cudaFree(s.ev_idx); cudaFree(s.loops);
    if (s.hblob) cudaFreeHost(s.hblob);
    s = ClSlot{};
}

int slot_ensure(ipc_handle* h, ClSlot& s, int L, int K) {
    const int PW = h->dim == 2 ? 5 : 7, NPQ = h->dim == 2 ? NPRE : se3::NP3, DQ = h->dim == 2 ? 3 : 6, LTW = h->dim == 2 ? 12 : CL3_LT;
    if (L > s.Lcap) {
        // the window buffers are O(n_poses): size them for the whole graph at once — cudaFree / cudaMalloc synchronise the device, and
        // a reallocation per growing window would serialise the side-by-side solves of the stream
        int cap = std::max(L, h->n);
        for (int q = 0; q < 2; ++q) { cudaFree(s.B[q].W); cudaFree(s.B[q].T); cudaFree(s.B[q].P); cudaFree(s.B[q].chi_e); s.B[q].W = s.B[q].T = s.B[q].P = s.B[q].chi_e = nullptr; }
        cudaFree(s.G); cudaFree(s.H); cudaFree(s.ev_ptr); s.G = s.H = nullptr; s.ev_ptr = nullptr; s.Lcap = 0;
        for (int q = 0; q < 2; ++q) {
            CUDA_TRY(cudaMalloc(&s.B[q].W, sizeof(double) * PW * (size_t)(cap + 1)));
            CUDA_TRY(cudaMalloc(&s.B[q].T, sizeof(double) * NPQ * (size_t)cap));
            CUDA_TRY(cudaMalloc(&s.B[q].P, sizeof(double) * NPQ * (size_t)(cap + 1)));
            CUDA_TRY(cudaMalloc(&s.B[q].chi_e, sizeof(double) * (size_t)cap));
        }
        CUDA_TRY(cudaMalloc(&s.G, sizeof(double) * DQ * (size_t)(cap + 1)));
        CUDA_TRY(cudaMalloc(&s.H, sizeof(double) * DQ * (size_t)(cap + 1)));
        CUDA_TRY(cudaMalloc(&s.ev_ptr, sizeof(int) * (size_t)(cap + 3)));
        s.Lcap = cap;
    }
    if (K > s.Kcap) {
        int cap = std::max(2 * K, 64);
        auto pad_of = [&](int k) { return ((size_t)DQ * k + CH_NB - 1) / CH_NB * CH_NB; };
        if (stream_smem_bytes((int)pad_of(cap)) > 226 * 1024) {
            cap = K;
            if (stream_smem_bytes((int)pad_of(cap)) > 226 * 1024)
                return fail(IPC_ERR_UNSUPPORTED, "cluster of " + std::to_string(K) + " loops exceeds the dense force-system solver (DESIGN.md)");
        }
        for (int q = 0; q < 2; ++q) { cudaFree(s.B[q].lt); s.B[q].lt = nullptr; }
        cudaFree(s.S); cudaFree(s.z); cudaFree(s.loops); cudaFree(s.lg); cudaFree(s.ev_idx);
        s.S = s.z = s.lg = nullptr; s.loops = nullptr; s.ev_idx = nullptr; s.Kcap = 0;
        for (int q = 0; q < 2; ++q) CUDA_TRY(cudaMalloc(&s.B[q].lt, sizeof(double) * LTW * (size_t)cap));
        const size_t np = pad_of(cap);
        CUDA_TRY(cudaMalloc(&s.S, sizeof(double) * (np + CH_NB) * np));
        CUDA_TRY(cudaMalloc(&s.z, sizeof(double) * np));
        CUDA_TRY(cudaMalloc(&s.lg, sizeof(double) * 2 * DQ * (size_t)cap));
        CUDA_TRY(cudaMalloc(&s.ev_idx, sizeof(int) * 2 * (size_t)cap));
        CUDA_TRY(cudaMalloc(&s.loops, std::max(sizeof(ClLoop), sizeof(ClLoop3)) * (size_t)cap));
        s.Kcap = cap;
    }
    return IPC_OK;
}

struct LoopRef { int from, to; const double* meas; const double* info; };

// Upload the sub-problem of one check into slot `si` (loop records, end-point events) and fill its kernel arguments.
// loops: the K loop edges, the candidate last. commit: see StreamArgs::commit.
int slot_prepare(ipc_handle* h, int si, int lo, int hi, const std::vector<LoopRef>& loops, double th, int iter_base, int commit, bool exact_iters,
                 StreamArgs& A, cudaStream_t st) {
    ClSlot& s = h->slots[si];
    const bool d2 = h->dim == 2;
    const int L = hi - lo, K = (int)loops.size(), DQ = d2 ? 3 : 6;
    int rc = slot_ensure(h, s, L, K);
    if (rc != IPC_OK) return rc;
    // PINNED host staging of this slot's uploads: copies from pageable memory block the calling thread, which would serialise
    // the side-by-side solves of the stream
    const size_t rec = d2 ? sizeof(ClLoop) : sizeof(ClLoop3);
    const size_t need = rec * K + sizeof(int) * ((size_t)L + 3 + 2 * (size_t)K);
    if (need > s.hblob_cap) {
        if (s.hblob) cudaFreeHost(s.hblob);
        s.hblob = nullptr; s.hblob_cap = 0;
        const size_t cap = need + need / 2 + 4096;
        CUDA_TRY(cudaMallocHost(&s.hblob, cap));
        s.hblob_cap = cap;
    }
    struct { unsigned char* p; unsigned char* data() { return p; } } blob{s.hblob};
    int* ptr = reinterpret_cast<int*>(blob.data() + rec * K);
    int* idx = ptr + (L + 3);
    std::fill(ptr, ptr + L + 3, 0);
    for (int l = 0; l < K; ++l) {
        const LoopRef& e = loops[l];
        const int jf = e.from - lo, jt = e.to - lo, a = std::min(jf, jt), b = std::max(jf, jt);
        if (d2) {
            ClLoop& o = reinterpret_cast<ClLoop*>(blob.data())[l];
            o.jf = jf; o.jt = jt; o.a = a; o.b = b;
            HostState::se2_edge_record(e.meas, e.info, 1.0, o.meas, o.D);
            HostState::inv_sym3_host(o.D, o.V);
        } else {
            ClLoop3& o = reinterpret_cast<ClLoop3*>(blob.data())[l];
            o.jf = jf; o.jt = jt; o.a = a; o.b = b;
            double r[49];
            if (!HostState::se3_edge_record(e.meas, e.info, 1.0, r)) return fail(IPC_ERR_ARG, "loop edge with a zero quaternion or a singular information matrix");
            for (int q = 0; q < 7; ++q) o.zinv[q] = r[q];
            for (int q = 0; q < 21; ++q) { o.Om[q] = r[7 + q]; o.V[q] = r[28 + q]; }
        }
        ++ptr[a + 1]; ++ptr[b + 1];
    }
    for (int j = 0; j <= L + 1; ++j) ptr[j + 1] += ptr[j];
    {   // loop end points by window position, loop order within a position (ClEvents)
        std::vector<int> fill(ptr, ptr + L + 2);
        for (int l = 0; l < K; ++l) {
            const int a = std::min(loops[l].from, loops[l].to) - lo, b = std::max(loops[l].from, loops[l].to) - lo;
            idx[fill[a]++] = (l << 1) | 1; idx[fill[b]++] = (l << 1);
        }
    }
    CUDA_TRY(cudaMemcpyAsync(s.loops, blob.data(), rec * K, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.ev_ptr, ptr, sizeof(int) * (L + 3), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.ev_idx, idx, sizeof(int) * 2 * (size_t)K, cudaMemcpyHostToDevice, st));
    A = StreamArgs{};
    A.dim = h->dim; A.lo = lo; A.L = L; A.K = K; A.n_poses = h->n; A.Lcap = s.Lcap;
    A.odom = h->cl_odom ? h->cl_odom : (d2 ? h->d_odom9 : h->d_odom49);
    A.odom_commit = d2 ? h->d_odom9 : h->d_odom49;
    A.pose = h->d_pose; A.loops = s.loops; A.ev = ClEvents{s.ev_ptr, s.ev_idx};
    A.B[0] = s.B[0]; A.B[1] = s.B[1]; A.G = s.G; A.H = s.H; A.lg = s.lg;
    A.n_pad = (DQ * K + CH_NB - 1) / CH_NB * CH_NB; A.ld = A.n_pad + CH_NB; A.S = s.S; A.z = s.z;
    A.res = h->cl_res + CL_NRES * si; A.stage3 = h->cl_stage; A.bar = h->cl_bar + 4 * si; A.ctl = reinterpret_cast<int*>(h->cl_bar + 4 * si + 1);
    A.th = th; A.max_iter = iter_base;
    if (!exact_iters && L + K > 100) A.max_iter *= 5;          // src/consensus_utils.cpp:12-13
    A.max_tries = h->max_tries; A.noise_eps = h->noise_eps; A.commit = commit; A.out = h->cl_out + 16 * si;
    A.prof = si == 0 ? h->cl_prof : nullptr;
    A.abort = nullptr;
    return IPC_OK;
}

// enqueue ONE check (slot si, a group of `group_size` CTAs) on `st`: barrier reset, arguments, cooperative launch, results to the
// pinned buffer. Several such launches on different streams run side by side (18 CTAs each on a 148-SM part).
int slot_launch(ipc_handle* h, int si, const StreamArgs& A, int group_size, cudaStream_t st, bool cooperative) {
    CUDA_TRY(cudaMemsetAsync(h->cl_bar + 4 * si, 0, sizeof(unsigned) * 4, st));
    h->slots[si].args_host = A;
    h->cl_hargs[si] = A;                                   // pinned
    CUDA_TRY(cudaMemcpyAsync(h->cl_args + si, h->cl_hargs + si, sizeof(StreamArgs), cudaMemcpyHostToDevice, st));
    const StreamArgs* d_args = h->cl_args + si;
    void* kargs[] = {&d_args, &group_size};
    const void* fn = h->dim == 2 ? (const void*)stream_check_kernel<2> : (const void*)stream_check_kernel<3>;
    // Cooperative launches do not overlap with one another; the side-by-side solves are plain launches instead. Their barriers need
    // every CTA of the group resident, which holds because the slots together never ask for more CTAs than the device has SMs
    // (stream_depth x group_size <= cl_grid, one CTA per SM) and a slot is only refilled after its previous kernel has finished.
    if (cooperative) CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(group_size), dim3(CL_NT), kargs, stream_smem_bytes(A.n_pad), st));
    else CUDA_TRY(cudaLaunchKernel(fn, dim3(group_size), dim3(CL_NT), kargs, stream_smem_bytes(A.n_pad), st));
    CUDA_TRY(cudaMemcpyAsync(h->cl_hout + 16 * si, h->cl_out + 16 * si, sizeof(double) * 16, cudaMemcpyDeviceToHost, st));
    return IPC_OK;
}

// one cooperative launch: group g of `group_size` CTAs solves args[g]; results of every slot come back through the pinned buffer
int launch_groups(ipc_handle* h, const std::vector<StreamArgs>& args, int group_size) {
    cudaStream_t st = h->stream;
    const int ng = (int)args.size();
    size_t smem = 0;
    for (const StreamArgs& a : args) smem = std::max(smem, stream_smem_bytes(a.n_pad));
    CUDA_TRY(cudaMemsetAsync(h->cl_bar, 0, sizeof(unsigned) * 4 * CL_MAX_SLOTS, st));
    for (int g = 0; g < ng; ++g) h->cl_hargs[g] = args[g];
    CUDA_TRY(cudaMemcpyAsync(h->cl_args, h->cl_hargs, sizeof(StreamArgs) * ng, cudaMemcpyHostToDevice, st));
    const StreamArgs* d_args = h->cl_args;
    void* kargs[] = {&d_args, &group_size};
    const void* fn = h->dim == 2 ? (const void*)stream_check_kernel<2> : (const void*)stream_check_kernel<3>;
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(ng * group_size), dim3(CL_NT), kargs, smem, st));
    CUDA_TRY(cudaMemcpyAsync(h->cl_hout, h->cl_out, sizeof(double) * 16 * ng, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int g = 0; g < ng; ++g) { const double* o = h->cl_hout + 16 * g; h->cl_n_fact += (long long)o[6]; h->cl_n_trial += (long long)o[7]; ++h->cl_n_checks; }
    return IPC_OK;
}

void info_from(const double* o, int L, int K, ipc_check_info* info) {
    if (!info) return;
    info->max_chi2 = o[1]; info->cand_chi2 = o[2]; info->sum_chi2 = o[3]; info->iterations = (int)o[4]; info->evals = (int)o[5]; info->window_len = L; info->n_loops = K;
}

// computeIndependentSubgraph (src/consensus.cpp:123-171) on the host mirror + the loop list of the sub-problem, candidate last
void cluster_of(ipc_handle* h, int from, int to, const double* meas, const double* info, int& lo, int& hi, std::vector<LoopRef>& loops) {
    std::vector<int> members;
    auto ext = h->hs.independent_subgraph(from, to, members);
    lo = ext.first; hi = ext.second;
    loops.clear();
    for (int m : members) { const HostEdge& e = h->hs.cns[m]; loops.push_back(LoopRef{e.from, e.to, e.meas.data(), e.info.data()}); }
    loops.push_back(LoopRef{from, to, meas, info});
}

}  // namespace

int ipc_agreement_check(ipc_handle* h, int from, int to, const double* meas, const double* info, int* accepted, ipc_check_info* out_info) {
    if (!h || !meas || !info || !accepted) return fail(IPC_ERR_ARG, "null argument");
    if (from < 0 || to < 0 || from >= h->n || to >= h->n || from == to) return fail(IPC_ERR_ARG, "invalid vertex ids");
    CUDA_TRY(cudaSetDevice(h->device));
    int lo, hi; std::vector<LoopRef> loops;
    cluster_of(h, from, to, meas, info, lo, hi, loops);
    const bool slow = loops.size() > 1;
    const double th = slow ? h->cfg.slow_reject_th : h->cfg.fast_reject_th;          // src/consensus.cpp:50-52
    const int ib = slow ? h->cfg.slow_reject_iter_base : h->cfg.fast_reject_iter_base;
    std::vector<StreamArgs> args(1);
    // accept = discard + push_back + propagateCurrentGuess (src/consensus.cpp:69-71), done by the kernel; a rejection leaves d_pose
    // untouched (restore): the solve works on a copy of the window
    int rc = slot_prepare(h, 0, lo, hi, loops, th, ib, /*commit=*/1, false, args[0], h->stream);
    if (rc != IPC_OK) return rc;
    rc = launch_groups(h, args, h->cl_grid);
    if (rc != IPC_OK) return rc;
    const bool ok = h->cl_hout[0] != 0.0;
    info_from(h->cl_hout, hi - lo, (int)loops.size(), out_info);
    *accepted = ok ? 1 : 0;
    if (ok) {
        HostEdge e; e.from = from; e.to = to; e.meas.assign(meas, meas + h->mw); e.info.assign(info, info + h->d * h->d);
        h->hs.cns.push_back(std::move(e));
    }
    return IPC_OK;
}

// The candidate loop of simulating_incremental_data (src/simulation.cpp:34-47) for n candidates in the given order, with exactly
// the sequential semantics of n ipc_agreement_check calls. Most candidates are rejected and a rejection leaves the IPC object
// untouched, so up to `stream_depth` candidates are in flight at once, each solved SPECULATIVELY against the current state by its
// own group of SMs (one cooperative launch per candidate, one CUDA stream per slot). Results are consumed strictly in order; a slot
// that finishes is refilled with the next candidate at once. An accept is committed (window stored, propagateCurrentGuess, consensus
// set grown); it invalidates every later solve in flight: those are told to give up (abort word) and are started again.
int ipc_agreement_check_stream(ipc_handle* h, int n, const int* from, const int* to, const double* meas, const double* info, int* accepted,
                               ipc_check_info* out_info) {
    if (!h || n < 0 || (n && (!from || !to || !meas || !info || !accepted))) return fail(IPC_ERR_ARG, "bad arguments");
    for (int i = 0; i < n; ++i)
        if (from[i] < 0 || to[i] < 0 || from[i] >= h->n || to[i] >= h->n || from[i] == to[i]) return fail(IPC_ERR_ARG, "candidate " + std::to_string(i) + " has invalid vertex ids");
    CUDA_TRY(cudaSetDevice(h->device));
    const int depth = std::max(1, std::min(h->stream_depth, CL_MAX_SLOTS));
    const int group_size = depth == 1 ? h->cl_grid : std::max(1, h->cl_grid / depth);
    const int mw = h->mw, dd = h->d * h->d;
    if (!h->slot_stream[0]) {
        for (int s = 0; s < CL_MAX_SLOTS; ++s) {
            CUDA_TRY(cudaStreamCreateWithFlags(&h->slot_stream[s], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&h->slot_ev[s], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&h->commit_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaHostAlloc(&h->h_abort, sizeof(int) * CL_MAX_SLOTS, cudaHostAllocMapped));
        std::fill(h->h_abort, h->h_abort + CL_MAX_SLOTS, 0);
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    struct Fl { int cand = -1; bool busy = false, done = false; int L = 0, K = 0; double t_launch = 0; };
    const bool trace = getenv("IPC_STREAM_TRACE") != nullptr;
    const auto t_origin = std::chrono::steady_clock::now();
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_origin).count(); };
    Fl fl[CL_MAX_SLOTS];
    int next_launch = 0, next_final = 0;
    auto fail_out = [&](int rc) { for (int s = 0; s < depth; ++s) cudaStreamSynchronize(h->slot_stream[s]); return rc; };
    while (next_final < n) {
        bool progressed = false;
        // refill free slots with the next candidates, solved against the CURRENT state
        for (int s = 0; s < depth && next_launch < n; ++s) {
            if (fl[s].busy || fl[s].done) continue;
            const int c = next_launch++;
            int lo, hi; std::vector<LoopRef> loops;
            cluster_of(h, from[c], to[c], meas + (size_t)mw * c, info + (size_t)dd * c, lo, hi, loops);
            const bool slow = loops.size() > 1;
            StreamArgs A;
            int rc = slot_prepare(h, s, lo, hi, loops, slow ? h->cfg.slow_reject_th : h->cfg.fast_reject_th,
                                  slow ? h->cfg.slow_reject_iter_base : h->cfg.fast_reject_iter_base, /*commit=*/0, false, A, h->slot_stream[s]);
            if (rc != IPC_OK) return fail_out(rc);
            A.abort = h->h_abort + s;
            rc = slot_launch(h, s, A, group_size, h->slot_stream[s], /*cooperative=*/depth == 1);
            if (rc != IPC_OK) return fail_out(rc);
            CUDA_TRY(cudaEventRecord(h->slot_ev[s], h->slot_stream[s]));
            fl[s].cand = c; fl[s].busy = true; fl[s].done = false; fl[s].L = hi - lo; fl[s].K = (int)loops.size(); fl[s].t_launch = now_ms();
            progressed = true;
        }
        for (int s = 0; s < depth; ++s) {
            if (!fl[s].busy) continue;
            const cudaError_t q = cudaEventQuery(h->slot_ev[s]);
            if (q == cudaSuccess) {
                fl[s].busy = false; fl[s].done = true; progressed = true;
                if (trace) std::fprintf(stderr, "slot %d cand %d K %d launched %.3f done %.3f ms\n", s, fl[s].cand, fl[s].K, fl[s].t_launch, now_ms());
            }
            else if (q != cudaErrorNotReady) return fail_out(fail(IPC_ERR_CUDA, std::string("stream slot: ") + cudaGetErrorString(q)));
        }
        // consume finished solves in candidate order
        for (;;) {
            int s = -1;
            for (int t = 0; t < depth; ++t) if (fl[t].done && fl[t].cand == next_final) s = t;
            if (s < 0) break;
            const double* o = h->cl_hout + 16 * s;
            const bool ok = o[0] != 0.0;
            accepted[next_final] = ok ? 1 : 0;
            if (out_info) info_from(o, fl[s].L, fl[s].K, out_info + next_final);
            h->cl_n_fact += (long long)o[6]; h->cl_n_trial += (long long)o[7]; ++h->cl_n_checks;
            fl[s].done = false; fl[s].cand = -1;
            const int c = next_final++;
            progressed = true;
            if (!ok) continue;
            // accept: every later solve in flight started from a state that no longer exists
            for (int t = 0; t < depth; ++t) if (fl[t].busy) h->h_abort[t] = 1;
            for (int t = 0; t < depth; ++t) if (fl[t].busy || fl[t].done) { CUDA_TRY(cudaStreamSynchronize(h->slot_stream[t])); ++h->cl_n_wasted; }
            for (int t = 0; t < depth; ++t) { h->h_abort[t] = 0; fl[t] = Fl{}; }
            next_launch = next_final;
            // commit slot s: store its window, propagateCurrentGuess, push_back (src/consensus.cpp:69-71)
            const StreamArgs& A = h->slots[s].args_host;
            const int cur = (int)o[8];
            if (h->dim == 2) stream_commit_kernel<2><<<1, CL_NT, 0, h->stream>>>(A.B[cur].W, h->d_pose, A.lo, A.L, h->d_odom9, h->n, h->cl_stage);
            else stream_commit_kernel<3><<<1, CL_NT, 0, h->stream>>>(A.B[cur].W, h->d_pose, A.lo, A.L, h->d_odom49, h->n, h->cl_stage);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(h->commit_ev, h->stream));
            for (int t = 0; t < depth; ++t) CUDA_TRY(cudaStreamWaitEvent(h->slot_stream[t], h->commit_ev, 0));
            HostEdge e; e.from = from[c]; e.to = to[c];
            e.meas.assign(meas + (size_t)mw * c, meas + (size_t)mw * (c + 1)); e.info.assign(info + (size_t)dd * c, info + (size_t)dd * (c + 1));
            h->hs.cns.push_back(std::move(e));
            break;
        }
        if (!progressed) std::this_thread::yield();
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return IPC_OK;
}

int ipc_final_optimize(ipc_handle* h, int max_iterations, double* chi2, int* iterations) {
    if (!h) return fail(IPC_ERR_ARG, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const bool d2 = h->dim == 2;
    if (d2 && !h->d_odom9_raw) {
        std::vector<double> rec;
        h->hs.build_odom_aos(false, h->n_pad, rec, /*raw=*/true);
        CUDA_TRY(cudaMalloc(&h->d_odom9_raw, rec.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(h->d_odom9_raw, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (!d2 && !h->d_odom49_raw) {
        std::vector<double> rec;
        if (!h->hs.build_odom_aos3(h->n_pad, rec, /*raw=*/true)) return fail(IPC_ERR_ARG, "singular odometry information");
        CUDA_TRY(cudaMalloc(&h->d_odom49_raw, rec.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(h->d_odom49_raw, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    // propagateGuess(0, N-1): vertex 0 at the origin and fixed, the rest dead-reckoned (src/simulation.cpp:50-53)
    if (d2) {
        cl_set_origin<<<1, 32, 0, h->stream>>>(h->d_pose);
        cl_dead_reckon<<<1, CL_NT, 0, h->stream>>>(h->d_odom9, 0, h->n, h->d_pose);
    } else {
        cl3_set_origin<<<1, 32, 0, h->stream>>>(h->d_pose);
        cl3_dead_reckon<<<1, CL_NT, 0, h->stream>>>(h->d_odom49, 0, h->n, h->d_pose, h->cl_stage);
    }
    CUDA_TRY(cudaGetLastError());
    const int K = (int)h->hs.cns.size();
    if (chi2) *chi2 = 0;
    if (iterations) *iterations = 0;
    if (K == 0) { CUDA_TRY(cudaStreamSynchronize(h->stream)); return IPC_OK; }
    const int lo = 0, hi = h->n - 1;
    std::vector<LoopRef> loops;
    for (const HostEdge& e : h->hs.cns) loops.push_back(LoopRef{e.from, e.to, e.meas.data(), e.info.data()});
    std::vector<StreamArgs> args(1);
    h->cl_odom = d2 ? h->d_odom9_raw : h->d_odom49_raw;        // odometry information / s_factor (src/simulation.cpp:55-56)
    int rc = slot_prepare(h, 0, lo, hi, loops, 0.0, max_iterations, /*commit=*/2, /*exact_iters=*/true, args[0], h->stream);
    h->cl_odom = nullptr;
    if (rc != IPC_OK) return rc;
    rc = launch_groups(h, args, h->cl_grid);
    if (rc != IPC_OK) return rc;
    ipc_check_info ci{};
    info_from(h->cl_hout, hi - lo, K, &ci);
    if (chi2) *chi2 = ci.sum_chi2;
    if (iterations) *iterations = ci.iterations;
    return IPC_OK;
}

int ipc_stream_profile(ipc_handle* h, double* out16, int reset) {
    if (!h || !out16) return fail(IPC_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long c[16];
    CUDA_TRY(cudaMemcpy(c, h->cl_prof, sizeof(c), cudaMemcpyDeviceToHost));
    int khz = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));
    for (int i = 0; i < 8; ++i) out16[i] = (double)c[i] / ((double)khz * 1e3);      // seconds at the nominal SM clock
    out16[8] = (double)h->cl_n_checks; out16[9] = (double)h->cl_n_fact; out16[10] = (double)h->cl_n_trial;
    for (int i = 11; i < 16; ++i) out16[i] = (double)c[i - 3] / ((double)khz * 1e3);    // factorisation sub-phases: diagonal block, panel solve, barrier, update, barrier
    if (reset) { CUDA_TRY(cudaMemset(h->cl_prof, 0, sizeof(c))); h->cl_n_checks = h->cl_n_fact = h->cl_n_trial = h->cl_n_wasted = 0; }
    return IPC_OK;
}

int ipc_get_poses(ipc_handle* h, double* out) {
    if (!h || !out) return fail(IPC_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * h->mw * (size_t)h->n));
    if (h->dim == 2) cl_export_poses<<<64, 256, 0, h->stream>>>(h->d_pose, h->n, d_out);
    else cl3_export_poses<<<64, 256, 0, h->stream>>>(h->d_pose, h->n, d_out);
    cudaError_t e = cudaMemcpyAsync(out, d_out, sizeof(double) * h->mw * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(IPC_ERR_CUDA, cudaGetErrorString(e));
    return IPC_OK;
}

namespace {
// Pairwise consistency matrix. With a communicator (ipc_comm_init) the solved checks are dealt round robin over the ranks
// (check c belongs to rank c % world: neighbours in the list have similar windows, so the deal is cost balanced), every rank
// solves its shard, ONE all-gather of the packed verdict words follows, and every rank assembles the same rows.
int consistency_matrix_impl(ipc_handle* h, uint32_t* rows_bits, int* order_out, int64_t* n_solved, bool sharded) {
    if (!h || !rows_bits) return fail(IPC_ERR_ARG, "null argument");
    if (!h->d_loops || h->n_loops <= 0) return fail(IPC_ERR_STATE, "no candidate table: call ipc_set_candidates first");
    if (sharded && !h->comm) return fail(IPC_ERR_STATE, "no communicator: call ipc_comm_init first");
    CUDA_TRY(cudaSetDevice(h->device));
    const int world = sharded ? h->comm->world() : 1, rank = sharded ? h->comm->rank() : 0;
    const int n = h->n_loops, words = (n + 31) / 32;
    // candidate order of src/simulation.cpp:26 (cmpTime), stable over file order
    std::vector<int> order(n), lo(n), hi(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::max(h->h_lfrom[a], h->h_lto[a]) < std::max(h->h_lfrom[b], h->h_lto[b]); });
    for (int i = 0; i < n; ++i) { lo[i] = std::min(h->h_lfrom[order[i]], h->h_lto[order[i]]); hi[i] = std::max(h->h_lfrom[order[i]], h->h_lto[order[i]]); }
    cudaStream_t st = h->stream;
    int *d_lo = nullptr, *d_hi = nullptr, *d_order = nullptr, *d_cnt = nullptr, *d_rowptr = nullptr, *d_total = nullptr;
    int *d_member = nullptr, *d_cand = nullptr, *d_pi = nullptr, *d_pj = nullptr, *d_work = nullptr, *d_lmember = nullptr, *d_lcand = nullptr;
    unsigned char *d_verdict = nullptr, *d_lverdict = nullptr; uint32_t *d_rows = nullptr, *d_gather = nullptr;
    auto cleanup = [&]() { cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_order); cudaFree(d_cnt); cudaFree(d_rowptr); cudaFree(d_total); cudaFree(d_member);
                           cudaFree(d_cand); cudaFree(d_pi); cudaFree(d_pj); cudaFree(d_work); cudaFree(d_verdict); cudaFree(d_rows);
                           cudaFree(d_lmember); cudaFree(d_lcand); cudaFree(d_lverdict); cudaFree(d_gather); };
#define MTRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { cleanup(); return fail(IPC_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); } } while (0)
    MTRY(cudaMalloc(&d_lo, sizeof(int) * n)); MTRY(cudaMalloc(&d_hi, sizeof(int) * n)); MTRY(cudaMalloc(&d_order, sizeof(int) * n));
    MTRY(cudaMalloc(&d_cnt, sizeof(int) * n)); MTRY(cudaMalloc(&d_rowptr, sizeof(int) * n)); MTRY(cudaMalloc(&d_total, sizeof(int)));
    MTRY(cudaMemcpyAsync(d_lo, lo.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    MTRY(cudaMemcpyAsync(d_hi, hi.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    MTRY(cudaMemcpyAsync(d_order, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    overlap_count<<<n, 256, 0, st>>>(d_lo, d_hi, n, d_cnt);
    row_offsets<<<1, 1024, 0, st>>>(d_cnt, n, n, d_rowptr, d_total);
    int total = 0;
    MTRY(cudaMemcpyAsync(&total, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
    MTRY(cudaStreamSynchronize(st));
    const int n_checks = total;           // n diagonal checks + overlapping pairs
    MTRY(cudaMalloc(&d_member, sizeof(int) * n_checks)); MTRY(cudaMalloc(&d_cand, sizeof(int) * n_checks));
    MTRY(cudaMalloc(&d_pi, sizeof(int) * n_checks)); MTRY(cudaMalloc(&d_pj, sizeof(int) * n_checks));
    MTRY(cudaMalloc(&d_verdict, n_checks));
    MTRY(cudaMalloc(&d_rows, sizeof(uint32_t) * (size_t)n * words));
    overlap_fill<<<n, 256, 0, st>>>(d_lo, d_hi, d_order, n, d_rowptr, d_member, d_cand, d_pi, d_pj);
    MTRY(cudaGetLastError());
    int rc = IPC_OK;
    if (world == 1) {
        MTRY(cudaMalloc(&d_work, sizeof(int) * (size_t)n_checks * NB));
        rc = enqueue_batch(h, n_checks, d_member, d_cand, d_work, n_checks, d_verdict, nullptr, nullptr, st);
        if (rc != IPC_OK) { cleanup(); return rc; }
    } else {
        const int n_local = (n_checks - rank + world - 1) / world;             // checks rank, rank + world, ...
        const int per = (n_checks + world - 1) / world;                         // the largest shard
        const size_t wpr = ((size_t)per + 31) / 32;                             // words per rank, equal on every rank
        MTRY(cudaMalloc(&d_lmember, sizeof(int) * std::max(n_local, 1))); MTRY(cudaMalloc(&d_lcand, sizeof(int) * std::max(n_local, 1)));
        MTRY(cudaMalloc(&d_lverdict, std::max(n_local, 1))); MTRY(cudaMalloc(&d_work, sizeof(int) * (size_t)std::max(n_local, 1) * NB));
        MTRY(cudaMalloc(&d_gather, sizeof(uint32_t) * wpr * world));
        MTRY(cudaMemsetAsync(d_gather + (size_t)rank * wpr, 0, sizeof(uint32_t) * wpr, st));
        if (n_local > 0) {
            shard_take<<<(n_local + 255) / 256, 256, 0, st>>>(d_member, d_cand, rank, world, n_local, d_lmember, d_lcand);
            MTRY(cudaGetLastError());
            rc = enqueue_batch(h, n_local, d_lmember, d_lcand, d_work, n_local, d_lverdict, d_gather + (size_t)rank * wpr, nullptr, st);
            if (rc != IPC_OK) { cleanup(); return rc; }
        }
        std::string err;
        if (!h->comm->all_gather_words(d_gather, wpr, st, err)) { cleanup(); return fail(IPC_ERR_CUDA, err); }
        shard_spread<<<(n_checks + 255) / 256, 256, 0, st>>>(d_gather, wpr, world, n_checks, d_verdict);
        MTRY(cudaGetLastError());
        h->last_launches += 2;
    }
    {
        const long long warps = (long long)n * words;
        const int threads = 256;
        const long long blocks = (warps * 32 + threads - 1) / threads;
        matrix_init<<<(unsigned)blocks, threads, 0, st>>>(d_lo, d_hi, d_verdict, n, words, d_rows);
        if (n_checks > n) matrix_scatter<<<(n_checks - n + 255) / 256, 256, 0, st>>>(d_verdict, d_pi, d_pj, n, n_checks, words, d_rows);
        MTRY(cudaGetLastError());
        h->last_launches += 5;
    }
    MTRY(cudaMemcpyAsync(rows_bits, d_rows, sizeof(uint32_t) * (size_t)n * words, cudaMemcpyDeviceToHost, st));
    MTRY(cudaStreamSynchronize(st));
#undef MTRY
    cleanup();
    if (order_out) std::copy(order.begin(), order.end(), order_out);
    if (n_solved) *n_solved = n_checks;
    return IPC_OK;
}
}  // namespace

int ipc_consistency_matrix(ipc_handle* h, uint32_t* rows_bits, int* order_out, int64_t* n_solved) {
    return consistency_matrix_impl(h, rows_bits, order_out, n_solved, false);
}
int ipc_consistency_matrix_sharded(ipc_handle* h, uint32_t* rows_bits, int* order_out, int64_t* n_solved) {
    return consistency_matrix_impl(h, rows_bits, order_out, n_solved, true);
}

// ---- multi-GPU: communicator + sharded batch (shard the independent checks, one all-gather) -------------
int ipc_comm_unique_id(unsigned char* id128) {
    if (!id128) return fail(IPC_ERR_ARG, "null argument");
    std::string err;
    if (!Comm::unique_id(id128, err)) return fail(IPC_ERR_CUDA, err);
    return IPC_OK;
}
int ipc_comm_init(ipc_handle* h, const unsigned char* id128, int rank, int world) {
    if (!h || !id128) return fail(IPC_ERR_ARG, "null argument");
    if (h->comm) return fail(IPC_ERR_STATE, "the handle already has a communicator");
    CUDA_TRY(cudaSetDevice(h->device));
    std::string err;
    h->comm = Comm::create(id128, rank, world, err);
    if (!h->comm) return fail(IPC_ERR_CUDA, err);
    return IPC_OK;
}
int ipc_comm_info(ipc_handle* h, int* rank, int* world, int64_t* n_collectives) {
    if (!h) return fail(IPC_ERR_ARG, "null handle");
    if (rank) *rank = h->comm ? h->comm->rank() : 0;
    if (world) *world = h->comm ? h->comm->world() : 1;
    if (n_collectives) *n_collectives = h->comm ? h->comm->n_collectives() : 0;
    return IPC_OK;
}

int ipc_check_batch_sharded_dev(ipc_handle* h, int n_local, const int* member_dev, const int* cand_dev, int words_per_rank,
                                uint32_t* out_bits_all_dev, void* stream) {
    if (!h || n_local < 0 || !out_bits_all_dev || words_per_rank < (n_local + 31) / 32) return fail(IPC_ERR_ARG, "bad arguments");
    if (!h->d_loops) return fail(IPC_ERR_STATE, "no candidate table: call ipc_set_candidates first");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rank = h->comm ? h->comm->rank() : 0;
    uint32_t* mine = out_bits_all_dev + (size_t)rank * words_per_rank;
    if (words_per_rank > n_local / 32) CUDA_TRY(cudaMemsetAsync(mine + n_local / 32, 0, sizeof(uint32_t) * (words_per_rank - n_local / 32), st));   // padding bits are zero
    if (n_local > 0) {
        int rc = ensure_batch_buffers(h, n_local);
        if (rc != IPC_OK) return rc;
        rc = enqueue_batch(h, n_local, member_dev, cand_dev, h->d_work, h->cap_checks, h->d_verdict, mine, nullptr, st);
        if (rc != IPC_OK) return rc;
    }
    if (h->comm && h->comm->world() > 1) {
        std::string err;
        if (!h->comm->all_gather_words(out_bits_all_dev, (size_t)words_per_rank, st, err)) return fail(IPC_ERR_CUDA, err);
    }
    return IPC_OK;
}

int ipc_check_batch_sharded(ipc_handle* h, int n_local, const int* member, const int* cand, int words_per_rank, uint32_t* out_bits_all) {
    if (!h || n_local < 0 || (n_local && (!member || !cand)) || !out_bits_all || words_per_rank < (n_local + 31) / 32) return fail(IPC_ERR_ARG, "bad arguments");
    if (!h->d_loops) return fail(IPC_ERR_STATE, "no candidate table: call ipc_set_candidates first");
    for (int i = 0; i < n_local; ++i)
        if (cand[i] < 0 || cand[i] >= h->n_loops || member[i] >= h->n_loops) return fail(IPC_ERR_ARG, "check " + std::to_string(i) + " indexes outside the candidate table");
    CUDA_TRY(cudaSetDevice(h->device));
    const int world = h->comm ? h->comm->world() : 1;
    int rc = ensure_batch_buffers(h, std::max(n_local, 1));
    if (rc != IPC_OK) return rc;
    const size_t all_words = (size_t)words_per_rank * world;
    if (all_words > h->gather_words) {
        cudaFree(h->d_gather); h->d_gather = nullptr; h->gather_words = 0;
        CUDA_TRY(cudaMalloc(&h->d_gather, sizeof(uint32_t) * all_words));
        h->gather_words = all_words;
    }
    cudaStream_t st = h->stream;
    if (n_local) {
        CUDA_TRY(cudaMemcpyAsync(h->d_member, member, sizeof(int) * n_local, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(h->d_cand, cand, sizeof(int) * n_local, cudaMemcpyHostToDevice, st));
    }
    rc = ipc_check_batch_sharded_dev(h, n_local, h->d_member, h->d_cand, words_per_rank, h->d_gather, st);
    if (rc != IPC_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_bits_all, h->d_gather, sizeof(uint32_t) * all_words, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return IPC_OK;
}

int ipc_greedy_consensus(ipc_handle* h, const uint32_t* rows_bits, int n, unsigned char* in_set) {
    if (!h || !rows_bits || !in_set || n < 0) return fail(IPC_ERR_ARG, "bad arguments");
    if (n == 0) return IPC_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int words = (n + 31) / 32;
    uint32_t *d_rows = nullptr, *d_S = nullptr; unsigned char* d_in = nullptr;
    auto cleanup = [&]() { cudaFree(d_rows); cudaFree(d_S); cudaFree(d_in); };
    cudaError_t e;
    if ((e = cudaMalloc(&d_rows, sizeof(uint32_t) * (size_t)n * words)) != cudaSuccess || (e = cudaMalloc(&d_S, sizeof(uint32_t) * words)) != cudaSuccess ||
        (e = cudaMalloc(&d_in, n)) != cudaSuccess) { cleanup(); return fail(IPC_ERR_CUDA, cudaGetErrorString(e)); }
    cudaMemcpyAsync(d_rows, rows_bits, sizeof(uint32_t) * (size_t)n * words, cudaMemcpyHostToDevice, h->stream);
    greedy_consensus<<<1, 512, 0, h->stream>>>(d_rows, n, words, d_S, d_in);
    cudaMemcpyAsync(in_set, d_in, n, cudaMemcpyDeviceToHost, h->stream);
    e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (e != cudaSuccess) return fail(IPC_ERR_CUDA, cudaGetErrorString(e));
    return IPC_OK;
}

}  // extern "C"
