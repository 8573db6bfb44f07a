// cpf_comm.cu -- the multi-GPU data plane of libcpf: one rank per GPU, particles partitioned by index, mesh replicated.
//
// The reference runs its whole CUDA path on the MPI master with ONE GPU: every rank sends its cells, points and
// velocities to the master (src/initCuda.H:207-270, src/advect.H:59-89 gatherList), the master expands and uploads, the
// other ranks idle.  Here every rank owns a context on its own GPU and the two per-step exchanges of the path run over
// NCCL (NVLink 5 / NVSwitch) on the context's stream, ordered with the sub-steps without host synchronisation:
//   * velocity field   rank `root` holds the solver's cell field -> ncclBroadcast into a device staging buffer ->
//                      repack kernel (cpf_update_velocity_bcast), or every rank contributes its own cells and the
//                      slices are all-gathered (cpf_update_velocity_slices: no gather-to-master at all);
//   * statistics       the fp64 statistics slot is summed over the ranks on the device before it crosses PCIe
//                      (cpf_stats_request).
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a process that already carries an NCCL (PyTorch) the same
// library instance is used, and a single-GPU run needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <vector>

#include "cpf_internal.h"

namespace cpf {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

static NcclApi g_nccl;
static std::string g_nccl_error;

static bool nccl_load()
{
    if (g_nccl.lib) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { g_nccl_error = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    NcclApi a;
    a.lib = h;
#define CPF_SYM(field, name)                                                                 \
    *(void **)(&a.field) = dlsym(h, name);                                                   \
    if (!a.field) { g_nccl_error = std::string("libnccl.so.2 lacks ") + name; dlclose(h); return false; }
    CPF_SYM(GetUniqueId, "ncclGetUniqueId")
    CPF_SYM(CommInitRank, "ncclCommInitRank")
    CPF_SYM(CommDestroy, "ncclCommDestroy")
    CPF_SYM(CommSplit, "ncclCommSplit")
    CPF_SYM(Broadcast, "ncclBroadcast")
    CPF_SYM(AllReduce, "ncclAllReduce")
    CPF_SYM(AllGather, "ncclAllGather")
    CPF_SYM(GroupStart, "ncclGroupStart")
    CPF_SYM(GroupEnd, "ncclGroupEnd")
    CPF_SYM(GetErrorString, "ncclGetErrorString")
    CPF_SYM(GetVersion, "ncclGetVersion")
#undef CPF_SYM
    g_nccl = a;
    return true;
}

struct CommState {
    ncclComm_t comm = nullptr;     // statistics (compute stream)
    ncclComm_t commField = nullptr; // velocity field (copy stream): its own communicator, so that the broadcast of step k+1
                                    // overlaps the sub-steps of step k instead of queueing behind them
    cudaEvent_t evCall = nullptr;
    int rank = 0, nranks = 1;
    double *d_stage = nullptr; // [nCells][3] broadcast / all-gather target in the solver's layout
    size_t stageBytes = 0;
    std::vector<long long> sliceOffset, sliceCount; // cpf_update_velocity_slices: cells owned by each rank
};

#define CPF_NCCL(ctx, call)                                                                              \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess)                                                                          \
            return fail(ctx, CPF_ERR_COMM, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

static int ensure_stage(cpf_context *ctx, size_t bytes)
{
    CommState *c = ctx->comm;
    if (bytes <= c->stageBytes) return CPF_OK;
    if (c->d_stage) { cudaStreamSynchronize(ctx->stream); cudaFree(c->d_stage); c->d_stage = nullptr; c->stageBytes = 0; }
    CPF_CUDA(ctx, cudaMalloc(&c->d_stage, bytes));
    c->stageBytes = bytes;
    return CPF_OK;
}

void comm_release(cpf_context *ctx)
{
    if (!ctx->comm) return;
    if (ctx->comm->commField && g_nccl.lib) g_nccl.CommDestroy(ctx->comm->commField);
    if (ctx->comm->comm && g_nccl.lib) g_nccl.CommDestroy(ctx->comm->comm);
    if (ctx->comm->evCall) cudaEventDestroy(ctx->comm->evCall);
    cudaFree(ctx->comm->d_stage);
    delete ctx->comm;
    ctx->comm = nullptr;
}

// in-place fp64 sum of a statistics slot over the ranks (no-op without a communicator)
int comm_reduce_stats(cpf_context *ctx, double *d_slot, int words)
{
    if (!ctx->comm || ctx->comm->nranks == 1) return CPF_OK;
    CPF_NCCL(ctx, g_nccl.AllReduce(d_slot, d_slot, (size_t)words, ncclDouble, ncclSum, ctx->comm->comm, ctx->stream));
    return CPF_OK;
}

// The staging buffer was last read by the repack kernel of the previous refresh (copy stream, in order), and the
// caller's data is ready in compute-stream order: the copy stream waits for "now" on the compute stream.
static int field_exchange_begin(cpf_context *ctx)
{
    CommState *c = ctx->comm;
    if (!c->evCall) CPF_CUDA(ctx, cudaEventCreateWithFlags(&c->evCall, cudaEventDisableTiming));
    CPF_CUDA(ctx, cudaEventRecord(c->evCall, ctx->stream));
    CPF_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, c->evCall, 0));
    return CPF_OK;
}

} // namespace cpf

using namespace cpf;

extern "C" {

int cpf_comm_unique_id(void *id, size_t bytes)
{
    if (!id || bytes < sizeof(ncclUniqueId)) return CPF_ERR_INVALID;
    if (!nccl_load()) return CPF_ERR_COMM;
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != ncclSuccess) return CPF_ERR_COMM;
    memset(id, 0, bytes);
    memcpy(id, &u, sizeof u);
    return CPF_OK;
}

int cpf_comm_init(cpf_context *ctx, const void *id, size_t bytes, int rank, int nranks)
{
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, CPF_ERR_INVALID, "cpf_comm_init: bad rank %d of %d", rank, nranks);
    comm_release(ctx);
    ctx->comm = new CommState;
    ctx->comm->rank = rank;
    ctx->comm->nranks = nranks;
    if (nranks == 1) return CPF_OK; // a communicator of one: every collective degenerates to a local copy, NCCL is not loaded
    if (!id || bytes < sizeof(ncclUniqueId)) { comm_release(ctx); return fail(ctx, CPF_ERR_INVALID, "cpf_comm_init: unique id of %zu bytes required", sizeof(ncclUniqueId)); }
    if (!nccl_load()) { comm_release(ctx); return fail(ctx, CPF_ERR_COMM, "%s", g_nccl_error.c_str()); }
    cudaSetDevice(ctx->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = g_nccl.CommInitRank(&ctx->comm->comm, nranks, u, rank);
    if (r != ncclSuccess) {
        const std::string msg = g_nccl.GetErrorString(r);
        ctx->comm->comm = nullptr;
        comm_release(ctx);
        return fail(ctx, CPF_ERR_COMM, "ncclCommInitRank failed: %s", msg.c_str());
    }
    r = g_nccl.CommSplit(ctx->comm->comm, 0, rank, &ctx->comm->commField, nullptr);
    if (r != ncclSuccess) {
        const std::string msg = g_nccl.GetErrorString(r);
        ctx->comm->commField = nullptr;
        comm_release(ctx);
        return fail(ctx, CPF_ERR_COMM, "ncclCommSplit failed: %s", msg.c_str());
    }
    return CPF_OK;
}

int cpf_comm_info(cpf_context *ctx, int *rank, int *nranks, int *ncclVersion)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
    if (nranks) *nranks = ctx->comm ? ctx->comm->nranks : 1;
    if (ncclVersion) { *ncclVersion = 0; if (g_nccl.lib) g_nccl.GetVersion(ncclVersion); }
    return CPF_OK;
}

int cpf_update_velocity_bcast(cpf_context *ctx, const double *U, int on_device, int root)
{
    if (!ctx || !ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_bcast: no mesh");
    if (!ctx->comm || ctx->comm->nranks == 1) {
        if (!U) return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_bcast: null field on the only rank");
        return cpf_update_velocity(ctx, U, on_device);
    }
    CommState *c = ctx->comm;
    if (root < 0 || root >= c->nranks) return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_bcast: bad root %d", root);
    if (c->rank == root && !U) return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_bcast: the root rank must pass the field");
    cudaSetDevice(ctx->device);
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->nCells;
    int rc = ensure_stage(ctx, bytes);
    if (rc) return rc;
    // upload (root), broadcast and repack run on the copy stream with their own communicator: they overlap the sub-steps
    // already enqueued on the compute stream (which read the other half of the double buffer)
    // a DEVICE source is ready in compute-stream order: the copy stream waits for "now" there.  A host source needs no
    // such wait (that wait would put the exchange of step k+1 behind the sub-steps of step k instead of beside them);
    // the staging buffer itself is only ever touched on the copy stream.
    if (c->rank == root && on_device == 1) { rc = field_exchange_begin(ctx); if (rc) return rc; } // on_device == 2: complete already
    const double *src = c->d_stage;
    if (c->rank == root) {
        if (on_device) src = U; // broadcast straight out of the caller's device buffer (ready in compute-stream order)
        else CPF_CUDA(ctx, cudaMemcpyAsync(c->d_stage, U, bytes, cudaMemcpyHostToDevice, ctx->copyStream));
    }
    CPF_NCCL(ctx, g_nccl.Broadcast(src, c->d_stage, 3 * (size_t)ctx->nCells, ncclDouble, root, c->commField, ctx->copyStream));
    return update_velocity_staged(ctx, c->d_stage);
}

int cpf_update_velocity_slices(cpf_context *ctx, long long cellOffset, long long nLocal, const double *Ulocal, int on_device)
{
    if (!ctx || !ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_slices: no mesh");
    if (cellOffset < 0 || nLocal < 0 || cellOffset + nLocal > ctx->nCells || (nLocal > 0 && !Ulocal))
        return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_slices: slice [%lld, %lld) outside the %lld cells", cellOffset, cellOffset + nLocal, ctx->nCells);
    cudaSetDevice(ctx->device);
    if (!ctx->comm) { ctx->comm = new CommState; }
    CommState *c = ctx->comm;
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->nCells;
    int rc = ensure_stage(ctx, bytes);
    if (rc) return rc;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (on_device == 1) { rc = field_exchange_begin(ctx); if (rc) return rc; }
    if (nLocal) CPF_CUDA(ctx, cudaMemcpyAsync(c->d_stage + 3 * cellOffset, Ulocal, sizeof(double) * 3 * (size_t)nLocal, kind, ctx->copyStream));
    if (c->nranks > 1) {
        // the slice table is exchanged once (and again whenever this rank's slice changes): 2 x 8 bytes per rank
        if ((int)c->sliceOffset.size() != c->nranks || c->sliceOffset[(size_t)c->rank] != cellOffset || c->sliceCount[(size_t)c->rank] != nLocal) {
            long long *d_tab = nullptr;
            CPF_CUDA(ctx, cudaMalloc(&d_tab, sizeof(long long) * 2 * (size_t)c->nranks));
            const long long mine[2] = { cellOffset, nLocal };
            CPF_CUDA(ctx, cudaMemcpyAsync(d_tab + 2 * c->rank, mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
            CPF_NCCL(ctx, g_nccl.AllGather(d_tab + 2 * c->rank, d_tab, 2, ncclInt64, c->comm, ctx->stream));
            std::vector<long long> tab(2 * (size_t)c->nranks);
            CPF_CUDA(ctx, cudaMemcpyAsync(tab.data(), d_tab, sizeof(long long) * tab.size(), cudaMemcpyDeviceToHost, ctx->stream));
            CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(d_tab);
            c->sliceOffset.assign((size_t)c->nranks, 0);
            c->sliceCount.assign((size_t)c->nranks, 0);
            long long covered = 0;
            for (int r = 0; r < c->nranks; ++r) { c->sliceOffset[(size_t)r] = tab[2 * (size_t)r]; c->sliceCount[(size_t)r] = tab[2 * (size_t)r + 1]; covered += tab[2 * (size_t)r + 1]; }
            if (covered != ctx->nCells) { c->sliceOffset.clear(); return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_slices: the ranks' slices cover %lld of %lld cells", covered, ctx->nCells); }
        }
        // variable-size all-gather: one broadcast per rank inside a group (NCCL fuses them into one launch)
        CPF_NCCL(ctx, g_nccl.GroupStart());
        for (int r = 0; r < c->nranks; ++r) {
            double *p = c->d_stage + 3 * c->sliceOffset[(size_t)r];
            ncclResult_t e = g_nccl.Broadcast(p, p, 3 * (size_t)c->sliceCount[(size_t)r], ncclDouble, r, c->commField, ctx->copyStream);
            if (e != ncclSuccess) { g_nccl.GroupEnd(); return fail(ctx, CPF_ERR_COMM, "ncclBroadcast (slice of rank %d) failed: %s", r, g_nccl.GetErrorString(e)); }
        }
        CPF_NCCL(ctx, g_nccl.GroupEnd());
    } else if (nLocal != ctx->nCells) {
        return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity_slices: a single rank must pass all %lld cells", ctx->nCells);
    }
    return update_velocity_staged(ctx, c->d_stage);
}

int cpf_set_particle_id_base(cpf_context *ctx, long long base)
{
    if (!ctx || base < 0) return fail(ctx, CPF_ERR_INVALID, "cpf_set_particle_id_base: negative base");
    if (ctx->rng_ready) return fail(ctx, CPF_ERR_INVALID, "cpf_set_particle_id_base: the generator states are already initialised");
    ctx->id_base = (unsigned long long)base;
    return CPF_OK;
}

} // extern "C"
