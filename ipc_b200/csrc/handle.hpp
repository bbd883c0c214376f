// handle.hpp — the handle behind include/ipc_b200.h: IPC<EDGE,VERTEX> state of
// /root/reference/include/ipc/consensus.hpp:23-32 kept as flat arrays, with the graph resident in HBM.
#pragma once
#include <string>
#include <vector>

#include "launch_common.hpp"
#include "host_state.hpp"
#include "stream_solver.cuh"
#include "comm.hpp"

using namespace ipcb;

constexpr int CL_MAX_SLOTS = 8;
struct ClSlot {                    // work buffers of one concurrent check of the sequential stream
    int Lcap = 0, Kcap = 0;        // capacities: window edges, loops
    ClBuffers B[2] = {};
    double *G = nullptr, *H = nullptr, *S = nullptr, *z = nullptr, *lg = nullptr;
    int *ev_ptr = nullptr, *ev_idx = nullptr;
    void* loops = nullptr;         // ClLoop (SE2) or ClLoop3 (SE3) records of the cluster
    unsigned char* hblob = nullptr; size_t hblob_cap = 0;   // pinned host staging of the slot's uploads
    StreamArgs args_host{};        // arguments of the slot's last launch
};

struct ipc_handle {
    int dim = 2, d = 3, mw = 3;
    int n = 0, n_pad = 0;
    int device = 0;
    int n_sm = 148;
    ipc_config cfg{};
    double noise_eps = 1e-13;         // DESIGN.md "Termination"; 0 = replay every retry like g2o
    int max_tries = 100;
    int speculate = 1;
    int early_accept = 0;
    int use_uniform = 1;              // allow the uniform-information kernels when the graph qualifies
    int sd_fuse = 2;                  // see CheckParams::sd_fuse
    Bucket buckets[NB];               // launch buckets (tunable: options bucket<i>_cap / bucket<i>_nt)
    HostState hs;                     // host mirror: odometry, consensus set (integer logic of consensus.cpp)
    // device graph
    double* d_odom9 = nullptr;        // AoS odometry records, general (9 doubles / edge)
    double* d_odom49 = nullptr;       // SE(3) AoS odometry records (Z^-1, Omega, Omega^-1: 49 doubles / edge)
    double* d_odom3 = nullptr;        // AoS odometry records, uniform isotropic information (3 doubles / edge); null if not applicable
    void* d_loops = nullptr;  int n_loops = 0;
    std::vector<int> h_lfrom, h_lto;  // host copy of candidate endpoints
    std::vector<double> h_lmeas, h_linfo;
    // batch work buffers (grown on demand)
    int cap_checks = 0;
    int *d_member = nullptr, *d_cand = nullptr, *d_work = nullptr, *d_counts = nullptr, *d_bucket_cap = nullptr;
    unsigned char* d_verdict = nullptr;
    uint32_t* d_bits = nullptr;
    ipc_check_info* d_info = nullptr;
    unsigned long long* d_stats = nullptr;
    double* d_scratch = nullptr; size_t scratch_doubles = 0;   // per-CTA scratch: one region per launch bucket (the bucket launches of a batch overlap)
    size_t scratch_off[NB] = {}, scratch_len[NB] = {};         // region of bucket b, in doubles
    int overlap_buckets = 1;                                   // option: 1 = every bucket on its own stream (fork / join around the batch), 0 = one after the other
    cudaStream_t bucket_stream[NB] = {};
    cudaEvent_t bucket_ev[NB] = {};
    cudaEvent_t fork_ev = nullptr;
    int last_launches = 0;
    cudaStream_t stream = nullptr;
    // ---- sequential stream (stateful agreementCheck): global pose state + cluster-solve work buffers
    double* d_pose = nullptr;         // AoS[5] x n: x y theta cos sin — the vertex estimates of the IPC object
    double* d_odom9_raw = nullptr;    // odometry records with the information as given (final optimisation only)
    const double* cl_odom = nullptr;  // records the cluster kernels read: d_odom9, or d_odom9_raw during ipc_final_optimize
    int cl_grid = 0;                                          // cooperative grid of the persistent solver: one CTA per SM
    int stream_depth = CL_MAX_SLOTS;                          // candidates solved side by side by ipc_agreement_check_stream (option stream_depth)
    std::vector<ClSlot> slots;                                // work buffers of the concurrent checks (slot 0: single checks, final optimisation)
    cudaStream_t slot_stream[CL_MAX_SLOTS] = {};              // one stream per slot: the speculative solves of the stream run side by side
    cudaEvent_t slot_ev[CL_MAX_SLOTS] = {};
    cudaEvent_t commit_ev = nullptr;
    int* h_abort = nullptr;                                   // host-mapped abort words, one per slot
    double* cl_res = nullptr;                                 // CL_NRES scalars per slot
    unsigned* cl_bar = nullptr;                               // per slot: group barrier counter + control words
    StreamArgs* cl_hargs = nullptr;                           // pinned copy of the kernel arguments
    StreamArgs* cl_args = nullptr;                            // kernel arguments, one record per group
    unsigned long long* cl_prof = nullptr;                    // phase cycle counters of slot 0's window CTA (ipc_stream_profile)
    long long cl_n_checks = 0, cl_n_fact = 0, cl_n_trial = 0, cl_n_wasted = 0;
    double *cl_out = nullptr, *cl_hout = nullptr;             // per-slot results (device / pinned host)
    double* cl_stage = nullptr;       // SE(3) dead-reckoning staging (CL_NT poses)
    double* d_odom49_raw = nullptr;   // SE(3) records with the information as given (final optimisation)
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // bracket the check kernels of the last batch (roofline timing)
    bool ev_valid = false;
    uint32_t* d_gather = nullptr; size_t gather_words = 0;   // [world][words_per_rank] verdict words of a sharded batch
    Comm* comm = nullptr;             // multi-GPU (comm.hpp): NCCL communicator created by ipc_comm_init, one handle per GPU
};
