// Multi-GPU plumbing of liboptex_b200: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference has no distributed code.  What shards, and how (DESIGN.md "Multi-GPU"):
//   cdf          PIXEL-sharded: rank g holds n_g rows of P and m_g rows of S (all C channels).  Rotations are per-row
//                work; the only cross-pixel quantities of histmatch.py:49-69 are the per-channel range (2 C words,
//                all-reduce MIN) and the two 256-bin histograms (2 C x 256 counts, all-reduce SUM).  Integer counts and
//                exact extrema: every rank builds the SAME tables as one GPU would - bit-identical results, 1 MB of
//                NVLink traffic per step at C = 512 instead of the 33.5 MB all-gather of a channel-sharded block.
//   chol/pca/sym PIXEL-sharded moments: column sums (C floats) and the centred Gram (C x C) are all-reduced (SUM);
//                the C x C chain runs redundantly on every rank, the application GEMM on the local rows.
//   sort         needs global ranks per channel: channel-sharded with an all-gather (optimaltextures_b200/parallel.py).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by the host process, e.g. torch's) so the
// library has no link-time dependency on it and a caller can hand in its own ncclComm_t.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

namespace optex {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static const NcclApi &nccl() {
    std::call_once(g_nccl_once, [] {
        // the instance the process already uses (torch links its bundled libnccl.so.2), else the system one
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        g_nccl.handle = h;
        g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
        g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
        g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
        g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
        g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
        g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce &&
                    g_nccl.AllGather && g_nccl.GetErrorString;
    });
    return g_nccl;
}

static int nccl_fail(ncclResult_t r, const char *what) {
    set_error("NCCL error %d (%s) at %s", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
    return OPTEX_ECUDA;
}
#define OPTEX_NCCL(call)                                    \
    do {                                                    \
        ncclResult_t _r = (call);                           \
        if (_r != ncclSuccess) return nccl_fail(_r, #call); \
    } while (0)

static int need_nccl() {
    if (nccl().ok) return OPTEX_OK;
    set_error("libnccl.so.2 could not be loaded (dlopen): the multi-GPU entry points need NCCL");
    return OPTEX_EUNSUPPORTED;
}

// reductions used by the sharded steps (api.cu, cov_match.cu)
int shard_allreduce_u32(const ShardComm *cm, uint32_t *buf, size_t count, bool take_min, cudaStream_t st) {
    if (!cm || cm->world <= 1) return OPTEX_OK;
    OPTEX_NCCL(nccl().AllReduce(buf, buf, count, ncclUint32, take_min ? ncclMin : ncclSum, (ncclComm_t)cm->comm, st));
    count_launch();
    return OPTEX_OK;
}
int shard_allreduce_f32_sum(const ShardComm *cm, float *buf, size_t count, cudaStream_t st) {
    if (!cm || cm->world <= 1) return OPTEX_OK;
    OPTEX_NCCL(nccl().AllReduce(buf, buf, count, ncclFloat32, ncclSum, (ncclComm_t)cm->comm, st));
    count_launch();
    return OPTEX_OK;
}
int shard_allgather_f32(const ShardComm *cm, const float *send, float *recv, size_t count_per_rank, cudaStream_t st) {
    if (!cm || cm->world <= 1) {
        if (send != recv) OPTEX_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * 4, cudaMemcpyDeviceToDevice, st));
        return OPTEX_OK;
    }
    OPTEX_NCCL(nccl().AllGather(send, recv, count_per_rank, ncclFloat32, (ncclComm_t)cm->comm, st));
    count_launch();
    return OPTEX_OK;
}

}  // namespace optex

using namespace optex;

extern "C" int optex_comm_unique_id(void *id_out, size_t id_bytes) {
    OPTEX_TRY(need_nccl());
    if (!id_out || id_bytes < sizeof(ncclUniqueId)) {
        set_error("optex_comm_unique_id: buffer of %zu bytes < %zu", id_bytes, sizeof(ncclUniqueId));
        return OPTEX_EINVAL;
    }
    ncclUniqueId id;
    OPTEX_NCCL(nccl().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return OPTEX_OK;
}

extern "C" size_t optex_comm_unique_id_bytes(void) { return sizeof(ncclUniqueId); }

extern "C" int optex_comm_init(const void *id, int rank, int world, optex_comm_t **out) {
    OPTEX_TRY(require_sm100());
    OPTEX_TRY(need_nccl());
    if (!id || !out || world < 1 || rank < 0 || rank >= world) {
        set_error("optex_comm_init: bad arguments (rank %d of %d)", rank, world);
        return OPTEX_EINVAL;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    OPTEX_NCCL(nccl().CommInitRank(&comm, world, uid, rank));
    ShardComm *cm = new ShardComm{(void *)comm, rank, world, 1};
    *out = (optex_comm_t *)cm;
    return OPTEX_OK;
}

extern "C" int optex_comm_adopt(void *nccl_comm, int rank, int world, optex_comm_t **out) {
    OPTEX_TRY(need_nccl());
    if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world) {
        set_error("optex_comm_adopt: bad arguments");
        return OPTEX_EINVAL;
    }
    *out = (optex_comm_t *)new ShardComm{nccl_comm, rank, world, 0};
    return OPTEX_OK;
}

extern "C" int optex_comm_destroy(optex_comm_t *comm) {
    ShardComm *cm = (ShardComm *)comm;
    if (!cm) return OPTEX_OK;
    if (cm->owned && cm->comm && nccl().ok) nccl().CommDestroy((ncclComm_t)cm->comm);
    delete cm;
    return OPTEX_OK;
}

extern "C" int optex_comm_rank(const optex_comm_t *comm) { return comm ? ((const ShardComm *)comm)->rank : -1; }
extern "C" int optex_comm_world(const optex_comm_t *comm) { return comm ? ((const ShardComm *)comm)->world : 0; }

// a plain collective on caller buffers, so host code above the C-ABI needs no second communicator
extern "C" int optex_comm_allgather_f32(optex_comm_t *comm, const float *send, float *recv, size_t count_per_rank,
                                        void *stream) {
    OPTEX_TRY(require_sm100());
    if (!comm || !send || !recv) {
        set_error("optex_comm_allgather_f32: NULL argument");
        return OPTEX_EINVAL;
    }
    return shard_allgather_f32((const ShardComm *)comm, send, recv, count_per_rank, (cudaStream_t)stream);
}
