// comm.hpp — the exchange step of the sharded check batch: ONE all-gather of packed verdict words over NCCL (NVLink 5 / NVSwitch).
//
// The reference has no parallelism beyond independent OS processes (/root/reference/bash/ipc_experiments_2D.sh:34-37); fast and
// pair checks are independent units, so the batch is dealt across the GPUs of one box with no data-path
// collective other than this gather. NCCL is bound at run time with dlopen("libnccl.so.2"): the library
// already mapped into the process (torch's bundled NCCL under Python, the system one under a C++ host) is reused, and
// libipc_b200.so itself has no link-time NCCL dependency — single-GPU hosts never touch it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace ipcb {

constexpr int COMM_ID_BYTES = 128;      // sizeof(ncclUniqueId)

class Comm {
public:
    // rank 0 creates the id and ships it to the other ranks by any means (torch.distributed broadcast, a file, a pipe ...)
    static bool unique_id(unsigned char* id128, std::string& err);
    // one communicator per (process | thread, device); the CUDA device must be current
    static Comm* create(const unsigned char* id128, int rank, int world, std::string& err);
    ~Comm();
    int rank() const { return rank_; }
    int world() const { return world_; }
    // in-place all-gather of `words` 32-bit words per rank: rank r's contribution sits at buf + r * words
    bool all_gather_words(uint32_t* buf, size_t words, cudaStream_t st, std::string& err);
    // number of all-gathers issued so far (bench.py reports it: the path has exactly one per batch)
    long long n_collectives() const { return n_coll_; }

private:
    Comm() = default;
    void* comm_ = nullptr;
    int rank_ = 0, world_ = 1;
    long long n_coll_ = 0;
};

}  // namespace ipcb
