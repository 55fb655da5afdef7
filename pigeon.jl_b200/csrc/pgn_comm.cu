// pgn_comm.cu — the one collective of the engine: the final gather of controls and per-QP statistics over NCCL (NVLink 5 / NVSwitch).
//
// The hot path shards the vehicle batch across GPUs with no inter-GPU traffic; a host that wants ONE contiguous result — the
// Julia deployment is a single process with one handle per GPU — calls pgn_gather after the loop.  Two ways to form the communicator:
//   pgn_comm_init_all(handles, n)                one process, one handle per GPU (ncclCommInitAll)
//   pgn_comm_unique_id + pgn_comm_init_rank      one process per GPU (torchrun-style launch); the id travels through the launcher
// NCCL is bound at run time with dlopen("libnccl.so.2"): a process that already loaded a libnccl (PyTorch bundles one) shares it, and a
// process that never gathers does not need NCCL at all.  Replaces nothing in the reference (it is single-vehicle); the loop that is sharded
// is reference src/model_predictive_control.jl:87-98.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "pgn_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;      // ncclSuccess = 0
enum { NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_FLOAT64 = 8 };      // ncclDataType_t values of nccl.h (stable across NCCL 2.x)

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    char err[256] = "";
};
Nccl g_nccl;


bool nccl_load() {
    Nccl& n = g_nccl;
    if (n.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) { snprintf(n.err, sizeof(n.err), "cannot load libnccl.so.2: %s", dlerror()); return false; }
#define SYM(field, name)                                                                                          \
    do {                                                                                                          \
        *(void**)(&n.field) = dlsym(n.lib, name);                                                                 \
        if (!n.field) { snprintf(n.err, sizeof(n.err), "libnccl lacks %s", name); dlclose(n.lib); n.lib = nullptr; return false; } \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommInitAll, "ncclCommInitAll"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
}

}  // namespace

extern "C" void pgn_set_error_(const char* msg);      // pgn_capi.cu: the thread-local message behind pgn_last_error()
#define SETERR(...)                                  \
    do {                                             \
        char b__[512];                               \
        snprintf(b__, sizeof(b__), __VA_ARGS__);     \
        pgn_set_error_(b__);                         \
    } while (0)

static int comm_fail(const char* what, int r) {
    SETERR("%s failed: %s", what, (g_nccl.GetErrorString && r) ? g_nccl.GetErrorString(r) : g_nccl.err);
    return PGN_ENCCL;
}
#define NC(call)                                             \
    do {                                                     \
        ncclResult_t r__ = (call);                           \
        if (r__ != 0) return comm_fail(#call, r__);          \
    } while (0)

static int alloc_gather(pgn_handle* h) {
    if (h->d_gath_c) return PGN_OK;
    const size_t n = (size_t)h->comm_size * h->B;
    void *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, n * 3 * 8 + 16) != cudaSuccess || cudaMalloc(&b, n * 2 * 4 + 16) != cudaSuccess) {
        SETERR("cudaMalloc of the gather buffers failed");
        return PGN_ENOMEM;
    }
    h->allocs.push_back(a); h->allocs.push_back(b);
    h->d_gath_c = (double*)a; h->d_gath_i = (int32_t*)b;
    return PGN_OK;
}

extern "C" {

int pgn_comm_unique_id(char* id) {
    if (!id) return PGN_EINVAL;
    if (!nccl_load()) return comm_fail("dlopen(libnccl)", 0);
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return PGN_OK;
}

int pgn_comm_init_rank(pgn_handle* h, int32_t nranks, int32_t rank, const char* id) {
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return PGN_EINVAL;
    if (!nccl_load()) return comm_fail("dlopen(libnccl)", 0);
    int prev = -1;
    cudaGetDevice(&prev); cudaSetDevice(h->device);
    pgn_comm_destroy(h);
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclComm_t c = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&c, nranks, u, rank);
    if (prev >= 0) cudaSetDevice(prev);
    if (r != 0) return comm_fail("ncclCommInitRank", r);
    h->comm = c; h->comm_rank = rank; h->comm_size = nranks;
    return PGN_OK;
}

int pgn_comm_init_all(pgn_handle* const* hs, int32_t n) {
    if (!hs || n < 1) return PGN_EINVAL;
    if (!nccl_load()) return comm_fail("dlopen(libnccl)", 0);
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        if (!hs[i]) return PGN_EINVAL;
        devs[i] = hs[i]->device;
        if (hs[i]->B != hs[0]->B) { SETERR("pgn_comm_init_all: every handle must hold the same batch size"); return PGN_EINVAL; }
        for (int j = 0; j < i; j++) if (devs[j] == devs[i]) { SETERR("pgn_comm_init_all: two handles on device %d", devs[i]); return PGN_EINVAL; }
        pgn_comm_destroy(hs[i]);
    }
    std::vector<ncclComm_t> comms(n);
    NC(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) { hs[i]->comm = comms[i]; hs[i]->comm_rank = i; hs[i]->comm_size = n; }
    return PGN_OK;
}

int pgn_comm_destroy(pgn_handle* h) {
    if (!h) return PGN_OK;
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr; h->comm_rank = 0; h->comm_size = 1;
    return PGN_OK;
}

// enqueue the all-gathers of one handle (inside a group when several handles of one thread take part)
static int gather_enqueue(pgn_handle* h) {
    const size_t B = h->B;
    // controls stay field-major per rank: [rank][3][B]; statistics [rank][B] each
    NC(g_nccl.AllGather(h->d_controls, h->d_gath_c, 3 * B, NCCL_FLOAT64, (ncclComm_t)h->comm, h->stream));
    NC(g_nccl.AllGather(h->d_iters, h->d_gath_i, B, NCCL_INT32, (ncclComm_t)h->comm, h->stream));
    NC(g_nccl.AllGather(h->d_status, h->d_gath_i + (size_t)h->comm_size * B, B, NCCL_INT32, (ncclComm_t)h->comm, h->stream));
    return PGN_OK;
}
static int gather_download(pgn_handle* h, double* controls, int32_t* iters, int32_t* status) {
    const size_t B = h->B, R = h->comm_size;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { SETERR("stream synchronisation after the gather failed: %s", cudaGetErrorString(cudaGetLastError())); return PGN_ECUDA; }
    if (controls) {
        std::vector<double> tmp(R * 3 * B);
        if (cudaMemcpy(tmp.data(), h->d_gath_c, tmp.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return PGN_ECUDA;
        for (size_t r = 0; r < R; r++) for (size_t v = 0; v < B; v++) for (size_t f = 0; f < 3; f++) controls[(r * B + v) * 3 + f] = tmp[(r * 3 + f) * B + v];
    }
    if (iters && cudaMemcpy(iters, h->d_gath_i, R * B * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return PGN_ECUDA;
    if (status && cudaMemcpy(status, h->d_gath_i + R * B, R * B * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return PGN_ECUDA;
    return PGN_OK;
}

int pgn_gather(pgn_handle* h, double* controls, int32_t* iters, int32_t* status) {
    if (!h) return PGN_EINVAL;
    if (!h->comm) { SETERR("pgn_gather: no communicator (call pgn_comm_init_rank or pgn_comm_init_all first)"); return PGN_ESTATE; }
    int prev = -1;
    cudaGetDevice(&prev); cudaSetDevice(h->device);
    int rc = alloc_gather(h);
    if (!rc) rc = gather_enqueue(h);
    if (!rc) rc = gather_download(h, controls, iters, status);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int pgn_gather_all(pgn_handle* const* hs, int32_t n, double* controls, int32_t* iters, int32_t* status) {
    if (!hs || n < 1) return PGN_EINVAL;
    for (int i = 0; i < n; i++) if (!hs[i] || !hs[i]->comm || hs[i]->comm_size != n) { SETERR("pgn_gather_all: the handles do not share a communicator of size %d", n); return PGN_ESTATE; }
    int prev = -1;
    cudaGetDevice(&prev);
    int rc = PGN_OK;
    for (int i = 0; i < n && !rc; i++) { cudaSetDevice(hs[i]->device); rc = alloc_gather(hs[i]); }
    if (!rc) {
        ncclResult_t r = g_nccl.GroupStart();       // one thread drives all ranks: the collectives must be grouped
        if (r != 0) rc = comm_fail("ncclGroupStart", r);
        for (int i = 0; i < n && !rc; i++) { cudaSetDevice(hs[i]->device); rc = gather_enqueue(hs[i]); }
        r = g_nccl.GroupEnd();
        if (!rc && r != 0) rc = comm_fail("ncclGroupEnd", r);
    }
    for (int i = 0; i < n && !rc; i++) {        // every rank holds the full result; the host arrays are filled from rank 0, the others are only drained
        cudaSetDevice(hs[i]->device);
        rc = i == 0 ? gather_download(hs[i], controls, iters, status) : gather_download(hs[i], nullptr, nullptr, nullptr);
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
