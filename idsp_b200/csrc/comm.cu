// comm.cu -- multi-GPU edges of the lane engine in the C ABI: idsp_b200_comm_* / idsp_scatter_lanes /
// idsp_gather_lanes (include/idsp_b200.h, SURVEY 8(b) "Signatures", 8(e)).
//
// Lanes never interact (dsp-process/src/compose.rs:472-475), so there is no collective inside the
// computation; the only exchange is handing contiguous lane blocks of a root-resident buffer to the
// ranks (one process per GPU) and collecting the results.  Both are grouped ncclSend / ncclRecv over
// NVLink 5 / NVSwitch on the ctx stream:
//   * lane-major  flat[l*frames + t]: a lane block is one contiguous range, sent / received in place
//     (zero staging copies on either side);
//   * frame-major flat[t*lanes + l]: a lane block is a strided [frames][hi-lo] window; the root packs
//     the remote blocks into a staging buffer (one 2-D device copy per peer) and sends them, receivers
//     get their block directly in its final contiguous form; the gather unpacks the same way.
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy already loaded into the process, e.g.
// torch's, or the system one), so libidsp_b200.so has no link-time dependency on it and single-GPU
// callers never touch it.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

// Minimal NCCL surface (nccl.h, stable since 2.7): declared here so the build needs no NCCL headers.
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;   // ncclSuccess = 0
typedef int ncclDataType_t; // ncclInt8 = 0
}
static_assert(sizeof(ncclUniqueId) == IDSP_COMM_ID_BYTES, "NCCL unique id size");

namespace {
struct Nccl {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok = false;
};

Nccl *nccl() {
    static Nccl n;  // C++11 static init: thread-safe
    static bool tried = false;
    if (tried) return n.ok ? &n : nullptr;
    tried = true;
    const char *names[] = {getenv("IDSP_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) break;
    }
    if (!n.handle) return nullptr;
#define SYM(field, name)                                                     \
    *(void **)(&n.field) = dlsym(n.handle, name);                            \
    if (!n.field) return nullptr;
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(Broadcast, "ncclBroadcast")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(GetVersion, "ncclGetVersion")
#undef SYM
    n.ok = true;
    return &n;
}
}  // namespace

struct idsp_comm {
    idsp_ctx *ctx;
    ncclComm_t comm;
    int nranks, rank;
    void *stage;         // frame-major pack / unpack staging (root only)
    size_t stage_bytes;
};

#define IDSP_NCCL(call)                                                                       \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != 0) {                                                                        \
            idsp_set_error("%s: %s failed: %s", __func__, #call, N->GetErrorString(r_));      \
            return IDSP_ENCCL;                                                                \
        }                                                                                     \
    } while (0)

static Nccl *need_nccl(const char *fn) {
    Nccl *N = nccl();
    if (!N) idsp_set_error("%s: NCCL (libnccl.so.2) could not be loaded: %s", fn, dlerror() ? dlerror() : "missing symbol");
    return N;
}

extern "C" int idsp_b200_comm_unique_id(unsigned char id[IDSP_COMM_ID_BYTES]) {
    if (!id) {
        idsp_set_error("idsp_b200_comm_unique_id: id is null");
        return IDSP_EINVAL;
    }
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    ncclUniqueId u;
    IDSP_NCCL(N->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return IDSP_OK;
}

extern "C" int idsp_b200_comm_init(idsp_ctx *ctx, int nranks, int rank, const unsigned char id[IDSP_COMM_ID_BYTES],
                                   idsp_comm **out) {
    int r = idsp_use_device(ctx);
    if (r) return r;
    IDSP_CHECK_ARG(out != nullptr, "out is null");
    *out = nullptr;
    IDSP_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "need 0 <= rank < nranks");
    IDSP_CHECK_ARG(id != nullptr || nranks == 1, "id is null");
    idsp_comm *c = new idsp_comm();
    c->ctx = ctx;
    c->comm = nullptr;
    c->nranks = nranks;
    c->rank = rank;
    c->stage = nullptr;
    c->stage_bytes = 0;
    if (nranks > 1) {
        Nccl *N = need_nccl(__func__);
        if (!N) {
            delete c;
            return IDSP_ENCCL;
        }
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclResult_t e = N->CommInitRank(&c->comm, nranks, u, rank);
        if (e != 0) {
            idsp_set_error("idsp_b200_comm_init: ncclCommInitRank failed: %s", N->GetErrorString(e));
            delete c;
            return IDSP_ENCCL;
        }
    }
    *out = c;
    return IDSP_OK;
}

extern "C" int idsp_b200_comm_free(idsp_comm *c) {
    if (!c) return IDSP_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->stage) cudaFree(c->stage);
    if (c->comm) {
        Nccl *N = nccl();
        if (N) N->CommDestroy(c->comm);
    }
    delete c;
    return IDSP_OK;
}

extern "C" int idsp_b200_comm_rank(const idsp_comm *c) { return c ? c->rank : -1; }
extern "C" int idsp_b200_comm_size(const idsp_comm *c) { return c ? c->nranks : 0; }
extern "C" int idsp_b200_nccl_version(void) {
    Nccl *N = nccl();
    int v = 0;
    if (!N || N->GetVersion(&v) != 0) return 0;
    return v;
}

// Contiguous lane block of `rank`: whole units of `align` lanes (a warp by default) except the last
// block, blocks cover [0, lanes) exactly and differ by at most one unit (idsp_b200/dist.py lane_block).
extern "C" int idsp_b200_lane_block(size_t lanes, int nranks, int rank, size_t align, size_t *lo, size_t *hi) {
    if (!lo || !hi || nranks < 1 || rank < 0 || rank >= nranks) {
        idsp_set_error("idsp_b200_lane_block: bad argument");
        return IDSP_EINVAL;
    }
    if (align == 0) align = 32;
    const size_t units = (lanes + align - 1) / align;
    size_t a = units * (size_t)rank / (size_t)nranks * align;
    size_t b = units * (size_t)(rank + 1) / (size_t)nranks * align;
    *lo = a < lanes ? a : lanes;
    *hi = b < lanes ? b : lanes;
    return IDSP_OK;
}

static int stage_reserve(idsp_comm *c, size_t bytes) {
    if (c->stage_bytes >= bytes) return IDSP_OK;
    if (c->stage) {
        IDSP_CUDA(cudaStreamSynchronize(c->ctx->stream));
        cudaFree(c->stage);
        c->stage = nullptr;
        c->stage_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&c->stage, bytes);
    if (e != cudaSuccess) {
        idsp_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        return IDSP_ENOMEM;
    }
    c->stage_bytes = bytes;
    return IDSP_OK;
}

#define EDGE_CHECK()                                                                           \
    do {                                                                                       \
        IDSP_CHECK_ARG(c != nullptr, "comm is null");                                          \
        int r_ = idsp_use_device(c->ctx);                                                      \
        if (r_) return r_;                                                                     \
        IDSP_CHECK_ARG(layout == IDSP_FRAME_MAJOR || layout == IDSP_LANE_MAJOR,                \
                       "layout must be 0 (frame-major) or 1 (lane-major)");                    \
        IDSP_CHECK_ARG(root >= 0 && root < c->nranks, "root out of range");                    \
        IDSP_CHECK_ARG(elem_bytes > 0, "elem_bytes must be > 0");                              \
        if (frames == 0 || lanes == 0) return IDSP_OK;                                         \
    } while (0)

// flat (all lanes, on root) -> part (lanes [lo, hi) of this rank, same layout, contiguous)
extern "C" int idsp_scatter_lanes(idsp_comm *c, const void *full, void *part, size_t frames, size_t lanes,
                                  size_t elem_bytes, int layout, int root) {
    EDGE_CHECK();
    cudaStream_t s = c->ctx->stream;
    size_t lo, hi;
    idsp_b200_lane_block(lanes, c->nranks, c->rank, 32, &lo, &hi);
    const size_t mine = (hi - lo) * frames * elem_bytes;
    IDSP_CHECK_ARG(part != nullptr || mine == 0, "part is null");
    const bool fm = layout == IDSP_FRAME_MAJOR;
    if (c->rank != root) {
        Nccl *N = need_nccl(__func__);
        if (!N) return IDSP_ENCCL;
        if (mine) IDSP_NCCL(N->Recv(part, mine, 0, root, c->comm, s));
        return IDSP_OK;
    }
    IDSP_CHECK_ARG(full != nullptr, "full is null on the root");
    const char *src = (const char *)full;
    // own block: one device copy (2-D when frame-major)
    if (mine) {
        if (fm)
            IDSP_CUDA(cudaMemcpy2DAsync(part, (hi - lo) * elem_bytes, src + lo * elem_bytes, lanes * elem_bytes,
                                        (hi - lo) * elem_bytes, frames, cudaMemcpyDeviceToDevice, s));
        else
            IDSP_CUDA(cudaMemcpyAsync(part, src + lo * frames * elem_bytes, mine, cudaMemcpyDeviceToDevice, s));
    }
    if (c->nranks == 1) return IDSP_OK;
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    if (fm) {
        // pack every remote block, then one grouped launch of sends
        int r = stage_reserve(c, (lanes - (hi - lo)) * frames * elem_bytes);
        if (r) return r;
        size_t off = 0;
        for (int p = 0; p < c->nranks; p++) {
            if (p == root) continue;
            size_t a, b;
            idsp_b200_lane_block(lanes, c->nranks, p, 32, &a, &b);
            if (b == a) continue;
            IDSP_CUDA(cudaMemcpy2DAsync((char *)c->stage + off, (b - a) * elem_bytes, src + a * elem_bytes,
                                        lanes * elem_bytes, (b - a) * elem_bytes, frames, cudaMemcpyDeviceToDevice, s));
            off += (b - a) * frames * elem_bytes;
        }
    }
    IDSP_NCCL(N->GroupStart());
    size_t off = 0;
    for (int p = 0; p < c->nranks; p++) {
        if (p == root) continue;
        size_t a, b;
        idsp_b200_lane_block(lanes, c->nranks, p, 32, &a, &b);
        const size_t n = (b - a) * frames * elem_bytes;
        if (!n) continue;
        const void *from = fm ? (const void *)((char *)c->stage + off) : (const void *)(src + a * frames * elem_bytes);
        ncclResult_t e = N->Send(from, n, 0, p, c->comm, s);
        if (e != 0) {
            N->GroupEnd();
            idsp_set_error("idsp_scatter_lanes: ncclSend failed: %s", N->GetErrorString(e));
            return IDSP_ENCCL;
        }
        off += n;
    }
    IDSP_NCCL(N->GroupEnd());
    return IDSP_OK;
}

// part (this rank's lane block) -> full (all lanes, on root); inverse of idsp_scatter_lanes
extern "C" int idsp_gather_lanes(idsp_comm *c, const void *part, void *full, size_t frames, size_t lanes,
                                 size_t elem_bytes, int layout, int root) {
    EDGE_CHECK();
    cudaStream_t s = c->ctx->stream;
    size_t lo, hi;
    idsp_b200_lane_block(lanes, c->nranks, c->rank, 32, &lo, &hi);
    const size_t mine = (hi - lo) * frames * elem_bytes;
    IDSP_CHECK_ARG(part != nullptr || mine == 0, "part is null");
    const bool fm = layout == IDSP_FRAME_MAJOR;
    if (c->rank != root) {
        Nccl *N = need_nccl(__func__);
        if (!N) return IDSP_ENCCL;
        if (mine) IDSP_NCCL(N->Send(part, mine, 0, root, c->comm, s));
        return IDSP_OK;
    }
    IDSP_CHECK_ARG(full != nullptr, "full is null on the root");
    char *dst = (char *)full;
    if (mine) {
        if (fm)
            IDSP_CUDA(cudaMemcpy2DAsync(dst + lo * elem_bytes, lanes * elem_bytes, part, (hi - lo) * elem_bytes,
                                        (hi - lo) * elem_bytes, frames, cudaMemcpyDeviceToDevice, s));
        else
            IDSP_CUDA(cudaMemcpyAsync(dst + lo * frames * elem_bytes, part, mine, cudaMemcpyDeviceToDevice, s));
    }
    if (c->nranks == 1) return IDSP_OK;
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    if (fm) {
        int r = stage_reserve(c, (lanes - (hi - lo)) * frames * elem_bytes);
        if (r) return r;
    }
    IDSP_NCCL(N->GroupStart());
    size_t off = 0;
    for (int p = 0; p < c->nranks; p++) {
        if (p == root) continue;
        size_t a, b;
        idsp_b200_lane_block(lanes, c->nranks, p, 32, &a, &b);
        const size_t n = (b - a) * frames * elem_bytes;
        if (!n) continue;
        void *to = fm ? (void *)((char *)c->stage + off) : (void *)(dst + a * frames * elem_bytes);
        ncclResult_t e = N->Recv(to, n, 0, p, c->comm, s);
        if (e != 0) {
            N->GroupEnd();
            idsp_set_error("idsp_gather_lanes: ncclRecv failed: %s", N->GetErrorString(e));
            return IDSP_ENCCL;
        }
        off += n;
    }
    IDSP_NCCL(N->GroupEnd());
    if (fm) {
        off = 0;
        for (int p = 0; p < c->nranks; p++) {
            if (p == root) continue;
            size_t a, b;
            idsp_b200_lane_block(lanes, c->nranks, p, 32, &a, &b);
            if (b == a) continue;
            IDSP_CUDA(cudaMemcpy2DAsync(dst + a * elem_bytes, lanes * elem_bytes, (char *)c->stage + off,
                                        (b - a) * elem_bytes, (b - a) * elem_bytes, frames, cudaMemcpyDeviceToDevice, s));
            off += (b - a) * frames * elem_bytes;
        }
    }
    return IDSP_OK;
}

// Small replicated data (coefficients, per-job parameters): root's `bytes` bytes to every rank, in place.
extern "C" int idsp_broadcast(idsp_comm *c, void *buf, size_t bytes, int root) {
    IDSP_CHECK_ARG(c != nullptr, "comm is null");
    int r = idsp_use_device(c->ctx);
    if (r) return r;
    IDSP_CHECK_ARG(root >= 0 && root < c->nranks, "root out of range");
    if (bytes == 0 || c->nranks == 1) return IDSP_OK;
    IDSP_CHECK_ARG(buf != nullptr, "buf is null");
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    IDSP_NCCL(N->Broadcast(buf, buf, bytes, 0, root, c->comm, c->ctx->stream));
    return IDSP_OK;
}

// ---------------------------------------------------------------- point-to-point building blocks
extern "C" int idsp_comm_group_begin(idsp_comm *c) {
    IDSP_CHECK_ARG(c != nullptr, "comm is null");
    if (c->nranks == 1) return IDSP_OK;
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    IDSP_NCCL(N->GroupStart());
    return IDSP_OK;
}
extern "C" int idsp_comm_group_end(idsp_comm *c) {
    IDSP_CHECK_ARG(c != nullptr, "comm is null");
    if (c->nranks == 1) return IDSP_OK;
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    IDSP_NCCL(N->GroupEnd());
    return IDSP_OK;
}
extern "C" int idsp_comm_send(idsp_comm *c, const void *buf, size_t bytes, int peer) {
    IDSP_CHECK_ARG(c != nullptr, "comm is null");
    int r = idsp_use_device(c->ctx);
    if (r) return r;
    IDSP_CHECK_ARG(peer >= 0 && peer < c->nranks && peer != c->rank, "peer out of range");
    if (bytes == 0) return IDSP_OK;
    IDSP_CHECK_ARG(buf != nullptr, "buf is null");
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    IDSP_NCCL(N->Send(buf, bytes, 0, peer, c->comm, c->ctx->stream));
    return IDSP_OK;
}
extern "C" int idsp_comm_recv(idsp_comm *c, void *buf, size_t bytes, int peer) {
    IDSP_CHECK_ARG(c != nullptr, "comm is null");
    int r = idsp_use_device(c->ctx);
    if (r) return r;
    IDSP_CHECK_ARG(peer >= 0 && peer < c->nranks && peer != c->rank, "peer out of range");
    if (bytes == 0) return IDSP_OK;
    IDSP_CHECK_ARG(buf != nullptr, "buf is null");
    Nccl *N = need_nccl(__func__);
    if (!N) return IDSP_ENCCL;
    IDSP_NCCL(N->Recv(buf, bytes, 0, peer, c->comm, c->ctx->stream));
    return IDSP_OK;
}
