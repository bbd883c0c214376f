// comm.cu — NCCL binding for the sharded check batch (see comm.hpp). Types come from <nccl.h>; the entry points are
// resolved with dlsym so that libipc_b200.so carries no DT_NEEDED on NCCL.
#include "comm.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

namespace ipcb {
namespace {

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi* api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        // RTLD_NOLOAD first: reuse the NCCL the process already mapped (torch's), else load by soname
        a.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!a.so) a.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.so) a.so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.so) { a.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
#define BIND(f) a.f = reinterpret_cast<decltype(a.f)>(dlsym(a.so, "nccl" #f)); if (!a.f) { a.err = "libnccl lacks nccl" #f; return; }
        BIND(GetUniqueId) BIND(CommInitRank) BIND(CommDestroy) BIND(AllGather) BIND(GetErrorString)
#undef BIND
    });
    return &a;
}

bool ok(ncclResult_t r, const char* what, std::string& err) {
    if (r == ncclSuccess) return true;
    err = std::string(what) + ": " + api()->GetErrorString(r);
    return false;
}

}  // namespace

bool Comm::unique_id(unsigned char* id128, std::string& err) {
    NcclApi* a = api();
    if (!a->err.empty()) { err = a->err; return false; }
    static_assert(sizeof(ncclUniqueId) == COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (!ok(a->GetUniqueId(&id), "ncclGetUniqueId", err)) return false;
    std::memcpy(id128, &id, COMM_ID_BYTES);
    return true;
}

Comm* Comm::create(const unsigned char* id128, int rank, int world, std::string& err) {
    NcclApi* a = api();
    if (!a->err.empty()) { err = a->err; return nullptr; }
    if (world < 1 || rank < 0 || rank >= world) { err = "bad rank / world"; return nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id128, COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    if (!ok(a->CommInitRank(&c, world, id, rank), "ncclCommInitRank", err)) return nullptr;
    Comm* o = new Comm();
    o->comm_ = c; o->rank_ = rank; o->world_ = world;
    return o;
}

Comm::~Comm() {
    if (comm_) api()->CommDestroy(static_cast<ncclComm_t>(comm_));
}

bool Comm::all_gather_words(uint32_t* buf, size_t words, cudaStream_t st, std::string& err) {
    ++n_coll_;
    return ok(api()->AllGather(buf + (size_t)rank_ * words, buf, words, ncclUint32, static_cast<ncclComm_t>(comm_), st), "ncclAllGather", err);
}

}  // namespace ipcb
