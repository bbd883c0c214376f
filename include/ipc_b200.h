/* ipc_b200.h — C ABI of the B200-native IPC hot path.
 *
 * The reference has no FFI: its boundary is the C++ class template IPC<EDGE,VERTEX>
 * (/root/reference/include/ipc/consensus.hpp:5-33) operating on g2o objects. This header is the
 * flat-array C ABI a binding of that class would call; every entry point cites the reference
 * interface it replaces. All pointers are plain host pointers unless the name says `_dev`.
 * Every function returns IPC_OK (0) or a negative error code; ipc_last_error() gives the text.
 * No entry point has a CPU fallback: without a CUDA device every compute call fails with
 * IPC_ERR_CUDA.
 *
 * Flat formats
 *   SE(2): measurement / pose = (x, y, theta), information = 3x3 row-major full symmetric.
 *   SE(3): measurement / pose = (x, y, z, qx, qy, qz, qw) (g2o EDGE_SE3:QUAT order, quaternion normalised on input),
 *          information = 6x6 row-major full symmetric (translation block first, then the quaternion-vector block).
 *   Vertex ids must be 0..n_poses-1 and odometry edge j must connect j -> j+1
 *   (the reference assumes both: src/consensus_utils.cpp:32-40, src/consensus.cpp:55).
 */
#ifndef IPC_B200_H
#define IPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipc_handle ipc_handle;

enum {
    IPC_OK = 0,
    IPC_ERR_ARG = -1,      /* bad argument (null pointer, bad id, bad dimension)            */
    IPC_ERR_CUDA = -2,     /* CUDA runtime error or no device                               */
    IPC_ERR_STATE = -3,    /* call not valid in the current state (e.g. no candidate table) */
    IPC_ERR_UNSUPPORTED = -4
};

/* The fields of struct Config the hot path reads (include/ipc/utils.hpp:22-38, consumed at
 * src/consensus.cpp:18-32). */
typedef struct ipc_config {
    double s_factor;
    double fast_reject_th;
    double slow_reject_th;
    int fast_reject_iter_base;
    int slow_reject_iter_base;
} ipc_config;

/* Per-check diagnostics (optional outputs). */
typedef struct ipc_check_info {
    double max_chi2;    /* max chi2 over every edge of the sub-problem after optimisation        */
    double cand_chi2;   /* chi2 of the candidate edge                                            */
    double sum_chi2;    /* sum over the sub-problem                                              */
    int iterations;     /* Dogleg iterations executed                                            */
    int evals;          /* step evaluations (tries)                                              */
    int window_len;     /* L = hi - lo                                                           */
    int n_loops;        /* K: loop edges in the sub-problem (candidate included)                 */
} ipc_check_info;

const char* ipc_last_error(void);
int ipc_device_count(void);

/* IPC<EDGE,VERTEX>::IPC(optimizer, cfg) — src/consensus.cpp:9-33: scales the odometry information
 * by s_factor (robustifyVoters, src/consensus_utils.cpp:123-130), dead-reckons every pose from the
 * origin (propagateGuess, :98-116) and snapshots. dim = 2 (EdgeSE2/VertexSE2) or 3 (EdgeSE3/VertexSE3).
 * odom_meas: [n_poses-1][3|7], odom_info: [n_poses-1][d*d]. device: CUDA ordinal. */
int ipc_create(int dim, int n_poses, const double* odom_meas, const double* odom_info,
               const ipc_config* cfg, int device, ipc_handle** out);
/* IPC::~IPC — src/consensus.cpp:35-40 */
void ipc_destroy(ipc_handle* h);

/* bool IPC::agreementCheck(EDGE*) — src/consensus.cpp:42-75. Stateful and sequential: cluster
 * discovery, fast/slow thresholds, Dogleg on the window, commit (consensus set grows, poses after the
 * window re-dead-reckoned) or rollback. *accepted = 1/0. info may be NULL. */
int ipc_agreement_check(ipc_handle* h, int from, int to, const double* meas, const double* info,
                        int* accepted, ipc_check_info* out_info);
/* The candidate loop of simulating_incremental_data — src/simulation.cpp:34-47: `for (edge : loops) ipc.agreementCheck(edge)` over
 * n candidates in the given order, with exactly the sequential semantics (and results) of n ipc_agreement_check calls. A rejection
 * leaves the IPC object untouched, so up to `stream_depth` (option, default 8) following candidates are solved speculatively side
 * by side, one group of SMs each; results are consumed in order up to the first accept, which is committed and invalidates the
 * rest of the round (those candidates are solved again against the new state). from/to: [n]; meas: [n][3|7]; info: [n][d*d];
 * accepted: [n] 1/0; out_info: [n] or NULL. */
int ipc_agreement_check_stream(ipc_handle* h, int n, const int* from, const int* to, const double* meas,
                               const double* info, int* accepted, ipc_check_info* out_info);
/* bool IPC::removeEdgeFromCnS(EDGE*) — src/consensus.cpp:77-98. *removed = 1/0. */
int ipc_remove_edge(ipc_handle* h, int from, int to, int* removed);
/* void IPC::addEdgeToCnS(EDGE*) — src/consensus.cpp:100-121. */
int ipc_add_edge(ipc_handle* h, int from, int to, const double* meas, const double* info);
/* getMaxConsensusSet() — include/ipc/consensus.hpp:16. from_to receives 2 ints per edge. */
int ipc_consensus_size(ipc_handle* h, int* n);
int ipc_get_consensus(ipc_handle* h, int* from_to, int capacity);
/* Current vertex estimates (what simulation.cpp:93-97 writes out): [n_poses][3|7]. */
int ipc_get_poses(ipc_handle* h, double* out);
/* The final full-graph optimisation of simulating_incremental_data — src/simulation.cpp:50-65: propagateGuess(0, N-1),
 * odometry information divided back by s_factor, Dogleg for at most max_iterations (1000 in the reference) on the odometry
 * chain + the consensus set. chi2 / iterations may be NULL. */
int ipc_final_optimize(ipc_handle* h, int max_iterations, double* chi2, int* iterations);

/* Diagnostics of the sequential stream (ipc_agreement_check / ipc_final_optimize run as one persistent kernel per call):
 * out16[0..7] = seconds (nominal SM clock) spent by the window CTA in: setup, assembly of the force system, Cholesky
 * factorisation, back-substitution, Gauss-Newton step, steepest-descent pass, trial states, commit; out16[8..10] = calls,
 * factorisations, trial states since the last reset. Non-reference. */
int ipc_stream_profile(ipc_handle* h, double* out16, int reset);

/* ---- batched, independent checks: the throughput path --------------------------------------
 * A check is one isAgreeingWithCurrentState test (src/consensus_utils.cpp:6-22) on a window that
 * starts from the dead-reckoned state of a fresh IPC object: with member < 0 it is exactly
 * `IPC ipc(...); ipc.agreementCheck(cand)` (fast path, K = 1); with member >= 0 it is
 * `IPC ipc(...); ipc.addEdgeToCnS(member); ipc.agreementCheck(cand)` (pair, K = 2 when the two
 * intervals overlap per src/consensus.cpp:157-159, otherwise the fast path on cand alone). */

/* Upload the loop-candidate table (device resident until replaced). */
int ipc_set_candidates(ipc_handle* h, int n_loops, const int* from, const int* to,
                       const double* meas, const double* info);
/* Host-buffer entry point (end-to-end: H2D of the check list, kernels, D2H of the results).
 * member/cand: [n_checks] indices into the candidate table. out_bits: ceil(n_checks/32) words,
 * bit c%32 of word c/32 = verdict of check c. out_info may be NULL. */
int ipc_check_batch(ipc_handle* h, int n_checks, const int* member, const int* cand,
                    uint32_t* out_bits, ipc_check_info* out_info);
/* Device-resident variant: all pointers are device pointers; kernels are enqueued on `stream`
 * (a cudaStream_t, may be NULL) and the call does not synchronise. ONE batch may be in flight per handle: the
 * work lists, verdict bytes and per-CTA scratch belong to the handle, so batches of one handle must be ordered — enqueue them on
 * the same stream (stream order serialises them; bench.py does that) or wait for the previous one before using another stream —
 * and member/cand must index the candidate table (not checked here — the host-buffer entry point validates). */
int ipc_check_batch_dev(ipc_handle* h, int n_checks, const int* member_dev, const int* cand_dev,
                        uint32_t* out_bits_dev, ipc_check_info* out_info_dev, void* stream);
/* Plan for ipc_check_batch_dev: the call sorts checks by window length into launch buckets on the
 * device; these report what the last batch processed (sum of window lengths L and loop counts K —
 * the algorithmic-bytes inputs of SURVEY.md §8(d)) and how many kernels it launched. */
int ipc_last_batch_stats(ipc_handle* h, int64_t* sum_L, int64_t* sum_K, int* n_launches);
/* Device time (CUDA events on the launching stream) spent in the check kernels of the last batch, i.e.
 * without the plan / pack kernels and without copies. Blocks until that batch has finished. */
int ipc_last_kernel_ms(ipc_handle* h, float* ms);

/* N_c x N_c pairwise consistency matrix over the candidate table, candidates in time order
 * (stable sort by max vertex id, src/simulation.cpp:26): diagonal = fast check, (i, j) i<j = pair
 * check of j against {i} when the intervals overlap, else diag(i) AND diag(j) (no solve).
 * rows_bits: [n_loops][ceil(n_loops/32)] words, symmetric. order_out (may be NULL): the time order
 * used, n_loops ints. n_solved (may be NULL): number of checks actually solved. */
int ipc_consistency_matrix(ipc_handle* h, uint32_t* rows_bits, int* order_out, int64_t* n_solved);

/* ---- multi-GPU: the check batch sharded over the GPUs of one box ----------------------------------
 * The reference is single-process (its only parallelism is independent OS processes,
 * bash/ipc_experiments_2D.sh:34-37); fast and pair checks are independent units, so a batch is dealt across
 * one handle per GPU (one process or thread each) and the ONLY exchange is one NCCL all-gather of the packed
 * verdict words. NCCL is bound at run time (dlopen of libnccl.so.2); single-GPU use never
 * touches it. */
#define IPC_COMM_ID_BYTES 128
/* Rank 0 creates the rendezvous id (an ncclUniqueId) and ships it to the other ranks by any means. */
int ipc_comm_unique_id(unsigned char id[IPC_COMM_ID_BYTES]);
/* Collective over the `world` handles (ncclCommInitRank on the handle's device). */
int ipc_comm_init(ipc_handle* h, const unsigned char id[IPC_COMM_ID_BYTES], int rank, int world);
/* rank / world of the handle (0 / 1 without a communicator) and the number of all-gathers issued so far. Any may be NULL. */
int ipc_comm_info(ipc_handle* h, int* rank, int* world, int64_t* n_collectives);
/* Sharded batch: this rank solves ITS n_local checks, then one in-place all-gather makes out_bits_all
 * [world][words_per_rank] identical on every rank: bit i%32 of word i/32 of row r = verdict of rank r's check i.
 * words_per_rank >= ceil(n_local/32) must be the same on every rank (pad to the largest shard). Without a
 * communicator the call is the single-rank batch (world = 1). Host buffers, end to end: */
int ipc_check_batch_sharded(ipc_handle* h, int n_local, const int* member, const int* cand, int words_per_rank,
                            uint32_t* out_bits_all);
/* ... and device-resident, enqueued on `stream` (kernels and the all-gather), no synchronisation. */
int ipc_check_batch_sharded_dev(ipc_handle* h, int n_local, const int* member_dev, const int* cand_dev,
                                int words_per_rank, uint32_t* out_bits_all_dev, void* stream);
/* ipc_consistency_matrix with the solved checks dealt round robin over the communicator's ranks (check c of the
 * time-ordered list belongs to rank c % world) and one all-gather; every rank receives the same rows. */
int ipc_consistency_matrix_sharded(ipc_handle* h, uint32_t* rows_bits, int* order_out, int64_t* n_solved);

/* Greedy consensus-set growth over a consistency matrix (row-AND + popcount): candidate k (in the
 * matrix's order) joins iff it is consistent with every current member. in_set: n_loops bytes. */
int ipc_greedy_consensus(ipc_handle* h, const uint32_t* rows_bits, int n, unsigned char* in_set);

/* Knobs (non-reference). noise_exit: 1 (default) stops the Dogleg retry loop once a rejected trial's own predicted gain is
 * below 1e-13 * chi2 (DESIGN.md "Termination"), 0 replays all 100 retries like g2o, any other value is used as the threshold.
 * early_accept: 1 lets verdict-only batches stop as soon as sum chi2 <= threshold (same bits). stream_depth (1..8): candidates
 * ipc_agreement_check_stream solves side by side. speculate, use_uniform,
 * max_tries, sd_fuse (0 / 1 / 2: when the steepest-descent pass replaces the norm pass; results identical),
 * bucket<i>_cap / bucket<i>_nt / bucket<i>_minb / bucket<i>_mode (launch shapes, i = 0..5; mode 0 = window state in shared memory,
 * 1 = streamed from global memory, 2 = + staged odometry): tuning, see DESIGN.md 3. cta_per_check: 1 selects the CTA-per-check launch
 * table (state in shared memory) on a uniform-information SE(2) graph instead of the default one-warp-per-check table. stage_odom.
 * overlap_buckets: 1 (default) runs the bucket launches of a batch side by side on internal streams (forked behind the caller's
 * stream, joined before the call's last kernel: stream order as seen by the caller is unchanged), 0 one after the other. */
int ipc_set_option(ipc_handle* h, const char* name, double value);

#ifdef __cplusplus
}
#endif
#endif /* IPC_B200_H */
