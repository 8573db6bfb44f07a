// cpf_api.cu -- the C ABI of libcpf (include/cpf.h).  Host-side orchestration only; every entry
// point validates, enqueues work on the context's stream and returns a status code.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "cpf_internal.h"

using namespace cpf;

namespace cpf {
void release_mesh(cpf_context *ctx);

int fail(cpf_context *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}



ParticleView particle_view(const cpf_context *ctx)
{
    ParticleView pv;
    const int a = ctx->pcur;
    pv.pos = ctx->d_pos[a]; pv.tet = ctx->d_tet[a]; pv.pid = ctx->d_pid[a]; pv.vel = ctx->d_vel[a]; pv.rng = ctx->d_rng[a];
    pv.n = ctx->n;
    return pv;
}

int ensure_scratch(cpf_context *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return CPF_OK;
    if (ctx->d_scratch) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0; }
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_scratch, bytes));
    ctx->scratch_bytes = bytes;
    return CPF_OK;
}

static void free_particles(cpf_context *ctx)
{
    for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->d_pos[b]); cudaFree(ctx->d_tet[b]); cudaFree(ctx->d_pid[b]); cudaFree(ctx->d_vel[b]); cudaFree(ctx->d_rng[b]);
        ctx->d_pos[b] = nullptr; ctx->d_tet[b] = nullptr; ctx->d_pid[b] = nullptr; ctx->d_vel[b] = nullptr; ctx->d_rng[b] = nullptr;
    }
    cudaFree(ctx->d_queue[0]); cudaFree(ctx->d_queue[1]); cudaFree(ctx->d_queue_count);
    ctx->d_queue[0] = ctx->d_queue[1] = nullptr; ctx->d_queue_count = nullptr;
    ctx->n = 0; ctx->pcur = 0; ctx->permuted = false; ctx->rng_ready = false; ctx->have_tets = false;
    ctx->statBaseValid = ctx->statScanQueued = false; // the next statistics request scans the new particle set
}

// solver layout [nCells][3] -> (ux,uy,uz,0) per cell: the hot kernels fetch a cell velocity with ONE 256-bit load
// (three 64-bit loads across two sectors were 20 % of the all-particles pass's L1 tag requests)
__global__ void k_pack_velocity(long long nCells, const double *__restrict__ U, double4 *__restrict__ out)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nCells) out[c] = make_double4(U[3 * c], U[3 * c + 1], U[3 * c + 2], 0.0);
}

// Continuous injection: every inactive particle (escaped through an ESCAPE patch, frozen outside the domain) is put back
// at a reproducible uniform position of the box, active, tet id unknown (-1: cpf_relocate_lost's BVH pass finds it).
// The position depends on (seed, sub-step index, GLOBAL particle id) only, not on the storage order.
__global__ void k_reseed_inactive(long long n, double4 *__restrict__ pos, int *__restrict__ tet, const int *__restrict__ pid,
                                  double4 *__restrict__ vel, double3 lo, double3 hi, unsigned long long key, unsigned long long idBase,
                                  unsigned long long *__restrict__ count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (i < n) {
        const double w = pos[i].w;
        if (w == 0.0) {
            hit = true;
            unsigned long long z = key + (idBase + (unsigned long long)pid[i]) * 0x9E3779B97F4A7C15ull;
            double u[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                z += 0x9E3779B97F4A7C15ull;
                unsigned long long x = z;
                x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
                x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
                x ^= x >> 31;
                u[c] = (double)(x >> 11) * (1.0 / 9007199254740992.0);
            }
            pos[i] = make_double4(lo.x + u[0] * (hi.x - lo.x), lo.y + u[1] * (hi.y - lo.y), lo.z + u[2] * (hi.z - lo.z), 1.0);
            tet[i] = -1;
            vel[i] = make_double4(0.0, 0.0, 0.0, -1.0);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(count, (unsigned long long)__popc(m));
}

__global__ void k_particle_cells(const MeshView m, long long n, const int *__restrict__ tet, const int *__restrict__ pid, int *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = tet[i];
    out[pid[i]] = t >= 0 ? tet_cell(m, t, ld_int4(m.tetv, t)) : -1;
}

__global__ void k_iota_fill(long long n, int *pid0, int *tet0)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { pid0[i] = (int)i; tet0[i] = -1; }
}

static int alloc_particles(cpf_context *ctx, long long n)
{
    free_particles(ctx);
    if (n <= 0) return CPF_OK;
    if (n >= (1ll << 31)) return fail(ctx, CPF_ERR_INVALID, "more than 2^31 particles per GPU are not supported");
    for (int b = 0; b < 2; ++b) {
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_pos[b], sizeof(double4) * (size_t)n));
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_tet[b], sizeof(int) * (size_t)n));
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_pid[b], sizeof(int) * (size_t)n));
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_vel[b], sizeof(double4) * (size_t)n));
        CPF_CUDA(ctx, cudaMemsetAsync(ctx->d_vel[b], 0, sizeof(double4) * (size_t)n, ctx->stream));
    }
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_queue[0], sizeof(int2) * (size_t)n));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_queue[1], sizeof(int2) * (size_t)n));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_queue_count, sizeof(unsigned) * 64));
    ctx->n = n;
    k_iota_fill<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->d_pid[0], ctx->d_tet[0]);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    return CPF_OK;
}

static inline unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
} // namespace cpf

static std::string g_create_error;

// streams, events and the counter block of a context; safe on a partially constructed one
static void destroy_handles(cpf_context *ctx)
{
    cudaFree(ctx->d_counters); ctx->d_counters = nullptr;
    cudaEvent_t *evs[] = { &ctx->ev0, &ctx->ev1, &ctx->evCopy, &ctx->evRead[0], &ctx->evRead[1] };
    for (cudaEvent_t *ev : evs) if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
    for (cudaEvent_t ev : ctx->profEvents) cudaEventDestroy(ev);
    ctx->profEvents.clear();
    if (ctx->ownStream) { cudaStreamDestroy(ctx->ownStream); ctx->ownStream = nullptr; }
    if (ctx->copyStream) { cudaStreamDestroy(ctx->copyStream); ctx->copyStream = nullptr; }
}

extern "C" {

int cpf_abi_version(void) { return CPF_ABI_VERSION; }

int cpf_device_count(int *n)
{
    if (!n) return CPF_ERR_INVALID;
    *n = 0;
    const cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess || *n == 0) { *n = 0; return CPF_ERR_NO_DEVICE; }
    return CPF_OK;
}

void cpf_default_config(cpf_config *cfg)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->device = 0;
    cfg->interp = CPF_INTERP_TET;       // src/initCuda.H:72
    cfg->locator = CPF_LOCATOR_CONVEX;  // etc/bashrc:1 RTX=false -> -DConvexPoly
    cfg->integrator = CPF_EULER;
    cfg->rng = CPF_RNG_XORWOW;          // usingBrownianMotion = true, src/initCuda.H:66
    cfg->reflect_wall = 1;              // src/initCuda.H:67
    cfg->path = CPF_PATH_FILTERED;
    cfg->sort_interval = 50;            // as the glue's sortInterval default (src/initCuda.H)
    cfg->fuse_substeps = 0;             // library default: 16 sub-steps per launch sequence (10 with the XORWOW stream)
    cfg->dt = 1e-4;                     // src/initCuda.H:55
    cfg->diffusion_coeff = 5.7e-6;      // src/initCuda.H:56
    cfg->seed = 1591593751ull;          // cuda/particles.cu:544
    cfg->save_interval = 10;            // src/initCuda.H:57
}

int cpf_create(const cpf_config *cfg, cpf_context **out)
{
    if (!out) return CPF_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libcpf has no CPU fallback";
        return CPF_ERR_NO_DEVICE;
    }
    cpf_context *ctx = new (std::nothrow) cpf_context;
    if (!ctx) return CPF_ERR_NOMEM;
    if (cfg) ctx->cfg = *cfg; else cpf_default_config(&ctx->cfg);
    ctx->device = ctx->cfg.device;
    if (ctx->device < 0 || ctx->device >= ndev) { g_create_error = "bad device ordinal"; delete ctx; return CPF_ERR_INVALID; }
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->evCopy, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->evRead[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->evRead[1], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * CNT_COUNT)) != cudaSuccess ||
        (e = cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * CNT_COUNT)) != cudaSuccess) {
        g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
        destroy_handles(ctx); // whatever was created before the failing call (every handle starts as nullptr)
        delete ctx;
        return CPF_ERR_CUDA;
    }
    ctx->stream = ctx->ownStream;
    *out = ctx;
    return CPF_OK;
}

int cpf_set_stream(cpf_context *ctx, void *cuda_stream)
{
    if (!ctx) return CPF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->ownStream;
    return CPF_OK;
}

int cpf_profile_enable(cpf_context *ctx, int enable)
{
    if (!ctx) return CPF_ERR_INVALID;
    ctx->profiling = enable != 0;
    ctx->profUsed = 0;
    return CPF_OK;
}

int cpf_profile_read(cpf_context *ctx, int *nLaunches, double *total_ms, double *max_ms)
{
    if (!ctx) return CPF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0.0, mx = 0.0;
    for (size_t q = 0; q + 1 < ctx->profUsed; q += 2) {
        float ms = 0.f;
        CPF_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->profEvents[q], ctx->profEvents[q + 1]));
        tot += ms;
        if (ms > mx) mx = ms;
    }
    if (nLaunches) *nLaunches = (int)(ctx->profUsed / 2);
    if (total_ms) *total_ms = tot;
    if (max_ms) *max_ms = mx;
    ctx->profUsed = 0;
    return CPF_OK;
}

int cpf_destroy(cpf_context *ctx)
{
    if (!ctx) return CPF_OK;
    cudaSetDevice(ctx->device);
    output_shutdown(ctx); // files handed to cpf_write_vtu_async are completed first
    cudaStreamSynchronize(ctx->stream);
    free_particles(ctx);
    release_mesh(ctx);
    comm_release(ctx);
    stats_release(ctx);
    cudaFree(ctx->d_sort_hist); cudaFree(ctx->d_scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    destroy_handles(ctx);
    delete ctx;
    return CPF_OK;
}

const char *cpf_last_error(const cpf_context *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int cpf_sync(cpf_context *ctx)
{
    if (!ctx) return CPF_ERR_INVALID;
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return output_wait(ctx);
}

int cpf_set_config(cpf_context *ctx, const cpf_config *cfg)
{
    if (!ctx || !cfg) return CPF_ERR_INVALID;
    if (cfg->device != ctx->cfg.device) return fail(ctx, CPF_ERR_INVALID, "the device cannot change after cpf_create");
    ctx->cfg = *cfg;
    return CPF_OK;
}

// ---------------------------------------------------------------------------------------------
// mesh
// ---------------------------------------------------------------------------------------------
static void bbox_of(cpf_context *ctx, long long n, const double *a, bool reset)
{
    if (reset) for (int k = 0; k < 3; ++k) { ctx->bbox_lo[k] = 1e300; ctx->bbox_hi[k] = -1e300; }
    for (long long i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            ctx->bbox_lo[k] = std::min(ctx->bbox_lo[k], a[3 * i + k]);
            ctx->bbox_hi[k] = std::max(ctx->bbox_hi[k], a[3 * i + k]);
        }
}

static int upload_patch_kinds(cpf_context *ctx, int nPatches, const int *patchKind)
{
    ctx->nPatches = nPatches > 0 ? nPatches : 1;
    std::vector<uint8_t> k((size_t)ctx->nPatches, (uint8_t)CPF_PATCH_REFLECT);
    if (patchKind)
        for (int p = 0; p < nPatches; ++p) {
            if (patchKind[p] != CPF_PATCH_REFLECT && patchKind[p] != CPF_PATCH_ESCAPE) return fail(ctx, CPF_ERR_INVALID, "patchKind[%d] = %d is not a cpf_patch_kind", p, patchKind[p]);
            k[(size_t)p] = (uint8_t)patchKind[p];
        }
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_patch_kind, k.size()));
    CPF_CUDA(ctx, cudaMemcpy(ctx->d_patch_kind, k.data(), k.size(), cudaMemcpyHostToDevice));
    const std::vector<double> gain((size_t)ctx->nPatches, 2.0); // specular everywhere, as the reference
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_patch_gain, sizeof(double) * gain.size()));
    CPF_CUDA(ctx, cudaMemcpy(ctx->d_patch_gain, gain.data(), sizeof(double) * gain.size(), cudaMemcpyHostToDevice));
    return CPF_OK;
}

extern "C" int cpf_set_patch_restitution(cpf_context *ctx, int nPatches, const double *e)
{
    if (!ctx || !ctx->have_mesh || !e) return fail(ctx, CPF_ERR_INVALID, "cpf_set_patch_restitution: no mesh or null table");
    if (nPatches != ctx->nPatches) return fail(ctx, CPF_ERR_INVALID, "cpf_set_patch_restitution: %d coefficients for %d patches", nPatches, ctx->nPatches);
    std::vector<double> gain((size_t)nPatches);
    for (int p = 0; p < nPatches; ++p) {
        if (!(e[p] > 0.0 && e[p] <= 1.0)) return fail(ctx, CPF_ERR_INVALID, "restitution[%d] = %g is outside (0, 1]", p, e[p]);
        gain[(size_t)p] = 1.0 + e[p];
    }
    cudaSetDevice(ctx->device);
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // kernels in flight read the table
    CPF_CUDA(ctx, cudaMemcpy(ctx->d_patch_gain, gain.data(), sizeof(double) * gain.size(), cudaMemcpyHostToDevice));
    return CPF_OK;
}

static int mesh_upload_poly(cpf_context *ctx, int nPoints, const double *points, int nFaces, const int *faceOffsets,
                            const int *faceVerts, const int *owner, int nInternal, const int *neighbour, int nCells,
                            const double *cellCentres, const int *tetBasePt, int nPatches, const int *patchStart, const int *patchKind);
// the ABI never throws: the host-side decomposition works in std::vector
int cpf_mesh_upload_poly(cpf_context *ctx, int nPoints, const double *points, int nFaces, const int *faceOffsets,
                         const int *faceVerts, const int *owner, int nInternal, const int *neighbour, int nCells,
                         const double *cellCentres, const int *tetBasePt, int nPatches, const int *patchStart, const int *patchKind)
{
    try {
        return mesh_upload_poly(ctx, nPoints, points, nFaces, faceOffsets, faceVerts, owner, nInternal, neighbour, nCells, cellCentres, tetBasePt,
                                nPatches, patchStart, patchKind);
    } catch (const std::bad_alloc &) {
        return fail(ctx, CPF_ERR_NOMEM, "cpf_mesh_upload_poly: out of host memory");
    } catch (const std::exception &e) {
        return fail(ctx, CPF_ERR_INVALID, "cpf_mesh_upload_poly: %s", e.what());
    }
}
static int mesh_upload_poly(cpf_context *ctx, int nPoints, const double *points, int nFaces, const int *faceOffsets,
                            const int *faceVerts, const int *owner, int nInternal, const int *neighbour, int nCells,
                            const double *cellCentres, const int *tetBasePt, int nPatches, const int *patchStart, const int *patchKind)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (!points || !faceOffsets || !faceVerts || !owner || (nInternal > 0 && !neighbour) || !cellCentres || nPoints <= 0 ||
        nFaces <= 0 || nCells <= 0 || nInternal < 0 || nInternal > nFaces)
        return fail(ctx, CPF_ERR_INVALID, "cpf_mesh_upload_poly: bad arguments");
    if (nPatches < 0 || (nPatches > 0 && !patchStart)) return fail(ctx, CPF_ERR_INVALID, "cpf_mesh_upload_poly: patchStart missing");
    for (int p = 0; p < nPatches; ++p)
        if (patchStart[p] < nInternal || patchStart[p] > patchStart[p + 1] || patchStart[p + 1] > nFaces)
            return fail(ctx, CPF_ERR_INVALID, "patchStart[%d..%d] = %d..%d is not a face range behind the %d internal faces", p, p + 1, patchStart[p], patchStart[p + 1], nInternal);
    cudaSetDevice(ctx->device);
    // --- the glue's decomposition (src/initCuda.H:86-110): cell by cell, face by face of the
    // cell (owned faces ascending, then neighbour faces ascending), one tet per face-fan triangle
    std::vector<int> start((size_t)nCells + 1, 0);
    for (int f = 0; f < nFaces; ++f) {
        if (owner[f] < 0 || owner[f] >= nCells) return fail(ctx, CPF_ERR_INVALID, "owner[%d] out of range", f);
        start[(size_t)owner[f] + 1]++;
    }
    for (int f = 0; f < nInternal; ++f) {
        if (neighbour[f] < 0 || neighbour[f] >= nCells) return fail(ctx, CPF_ERR_INVALID, "neighbour[%d] out of range", f);
        start[(size_t)neighbour[f] + 1]++;
    }
    for (int c = 0; c < nCells; ++c) start[(size_t)c + 1] += start[(size_t)c];
    std::vector<int> cellFaces((size_t)start[(size_t)nCells]), fill((size_t)nCells, 0);
    for (int f = 0; f < nFaces; ++f) { const int c = owner[f]; cellFaces[(size_t)(start[(size_t)c] + fill[(size_t)c]++)] = f; }
    for (int f = 0; f < nInternal; ++f) { const int c = neighbour[f]; cellFaces[(size_t)(start[(size_t)c] + fill[(size_t)c]++)] = f; }
    std::vector<int> facePatch;
    if (nPatches > 0 && patchStart) {
        facePatch.assign((size_t)nFaces, -1);
        for (int p = 0; p < nPatches; ++p)
            for (int f = patchStart[p]; f < patchStart[p + 1] && f < nFaces; ++f) facePatch[(size_t)f] = p;
    }
    long long nTets = 0;
    for (int f = 0; f < nFaces; ++f) {
        const int sz = faceOffsets[f + 1] - faceOffsets[f];
        if (sz < 3) return fail(ctx, CPF_ERR_INVALID, "face %d has fewer than 3 points", f);
        if (tetBasePt && (tetBasePt[f] < 0 || tetBasePt[f] >= sz)) return fail(ctx, CPF_ERR_INVALID, "tetBasePt[%d] out of range", f);
        for (int q = faceOffsets[f]; q < faceOffsets[f + 1]; ++q)
            if (faceVerts[q] < 0 || faceVerts[q] >= nPoints) return fail(ctx, CPF_ERR_INVALID, "face %d references point %d out of range", f, faceVerts[q]);
        nTets += (long long)(sz - 2) * (f < nInternal ? 2 : 1);
    }
    std::vector<int> tets((size_t)nTets * 4), tetPatch((size_t)nTets, -1);
    long long t = 0;
    for (int c = 0; c < nCells; ++c)
        for (int q = start[(size_t)c]; q < start[(size_t)c + 1]; ++q) {
            const int f = cellFaces[(size_t)q];
            const int *F = faceVerts + faceOffsets[f];
            const int sz = faceOffsets[f + 1] - faceOffsets[f];
            const int base = tetBasePt ? tetBasePt[f] : 0;
            const bool own = owner[f] == c;
            for (int tp = 1; tp <= sz - 2; ++tp) {
                int pa = (tp + base) % sz, pb = (pa + 1) % sz;
                if (!own) std::swap(pa, pb);
                int *o = &tets[(size_t)t * 4];
                o[0] = nPoints + c; o[1] = F[base]; o[2] = F[pa]; o[3] = F[pb];
                if (!facePatch.empty()) tetPatch[(size_t)t] = facePatch[(size_t)f];
                else tetPatch[(size_t)t] = f >= nInternal ? 0 : -1;
                ++t;
            }
        }
    // vertex array = [points..., cell centres...]  (src/initCuda.H:112-124)
    std::vector<double> pos(((size_t)nPoints + (size_t)nCells) * 3);
    memcpy(pos.data(), points, sizeof(double) * 3 * (size_t)nPoints);
    memcpy(pos.data() + 3 * (size_t)nPoints, cellCentres, sizeof(double) * 3 * (size_t)nCells);
    bbox_of(ctx, nPoints, points, true);
    int rc = build_device_mesh(ctx, (long long)nPoints + nCells, pos.data(), nTets, tets.data(), nullptr, tetPatch.data(), nCells,
                               nPoints, true);
    if (rc) return rc;
    cudaFree(ctx->d_pc_off); cudaFree(ctx->d_pc_cells); ctx->d_pc_off = ctx->d_pc_cells = nullptr;
    if (ctx->cfg.interp == CPF_INTERP_VERTEX) {
        // point -> cells CSR (cells ascending, unique) for the point-value interpolation
        std::vector<long long> pairs;
        pairs.reserve((size_t)faceOffsets[nFaces] * 2);
        for (int f = 0; f < nFaces; ++f)
            for (int q = faceOffsets[f]; q < faceOffsets[f + 1]; ++q) {
                pairs.push_back(((long long)faceVerts[q] << 32) | (unsigned)owner[f]);
                if (f < nInternal) pairs.push_back(((long long)faceVerts[q] << 32) | (unsigned)neighbour[f]);
            }
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        std::vector<int> off((size_t)nPoints + 1, 0), cl(pairs.size());
        for (size_t k = 0; k < pairs.size(); ++k) { off[(size_t)(pairs[k] >> 32) + 1]++; cl[k] = (int)(pairs[k] & 0xffffffffll); }
        for (int q = 0; q < nPoints; ++q) off[(size_t)q + 1] += off[(size_t)q];
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_pc_off, sizeof(int) * off.size()));
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_pc_cells, sizeof(int) * std::max<size_t>(cl.size(), 1)));
        CPF_CUDA(ctx, cudaMemcpy(ctx->d_pc_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
        CPF_CUDA(ctx, cudaMemcpy(ctx->d_pc_cells, cl.data(), sizeof(int) * cl.size(), cudaMemcpyHostToDevice));
    }
    return upload_patch_kinds(ctx, nPatches, patchKind);
}

static int mesh_upload_tets(cpf_context *ctx, int nVerts, const double *positions, long long nTets, const int *tetVerts, const int *tetCell, int nCells);
int cpf_mesh_upload_tets(cpf_context *ctx, int nVerts, const double *positions, long long nTets, const int *tetVerts,
                         const int *tetCell, int nCells)
{
    try {
        return mesh_upload_tets(ctx, nVerts, positions, nTets, tetVerts, tetCell, nCells);
    } catch (const std::bad_alloc &) {
        return fail(ctx, CPF_ERR_NOMEM, "cpf_mesh_upload_tets: out of host memory");
    } catch (const std::exception &e) {
        return fail(ctx, CPF_ERR_INVALID, "cpf_mesh_upload_tets: %s", e.what());
    }
}
static int mesh_upload_tets(cpf_context *ctx, int nVerts, const double *positions, long long nTets, const int *tetVerts, const int *tetCell, int nCells)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (!positions || !tetVerts || nVerts <= 0 || nTets <= 0) return fail(ctx, CPF_ERR_INVALID, "cpf_mesh_upload_tets: bad arguments");
    cudaSetDevice(ctx->device);
    if (!tetCell) nCells = (int)nTets;
    else {
        if (nCells <= 0) return fail(ctx, CPF_ERR_INVALID, "cpf_mesh_upload_tets: nCells must be positive when tetCell is given");
        for (long long t = 0; t < nTets; ++t)
            if (tetCell[t] < 0 || tetCell[t] >= nCells) return fail(ctx, CPF_ERR_INVALID, "tetCell[%lld] = %d out of range", t, tetCell[t]);
    }
    bbox_of(ctx, nVerts, positions, true);
    int rc = build_device_mesh(ctx, nVerts, positions, nTets, tetVerts, tetCell, nullptr, nCells, 0, false);
    if (rc) return rc;
    return upload_patch_kinds(ctx, 1, nullptr);
}

int cpf_mesh_info(cpf_context *ctx, long long *nVerts, long long *nTets, long long *nCells, long long *nBoundaryFaces)
{
    if (!ctx || !ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "no mesh uploaded");
    if (nVerts) *nVerts = ctx->nVerts;
    if (nTets) *nTets = ctx->nTets;
    if (nCells) *nCells = ctx->nCells;
    if (nBoundaryFaces) *nBoundaryFaces = ctx->nBoundaryFaces;
    return CPF_OK;
}

int cpf_mesh_download_tets(cpf_context *ctx, int *tetVerts, int *tetCell)
{
    if (!ctx || !ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "no mesh uploaded");
    cudaSetDevice(ctx->device);
    std::vector<int4> sv((size_t)ctx->nTets);
    std::vector<uint16_t> code((size_t)ctx->nTets);
    CPF_CUDA(ctx, cudaMemcpy(sv.data(), ctx->d_tetv, sizeof(int4) * sv.size(), cudaMemcpyDeviceToHost));
    CPF_CUDA(ctx, cudaMemcpy(code.data(), ctx->d_tetcode, sizeof(uint16_t) * code.size(), cudaMemcpyDeviceToHost));
    std::vector<int> cells;
    if (tetCell && !ctx->cellFromVertex) {
        CPF_CUDA(ctx, cudaMemcpy(tetCell, ctx->d_tetcell, sizeof(int) * (size_t)ctx->nTets, cudaMemcpyDeviceToHost));
    }
    for (long long t = 0; t < ctx->nTets; ++t) {
        const int s[4] = { sv[(size_t)t].x, sv[(size_t)t].y, sv[(size_t)t].z, sv[(size_t)t].w };
        if (tetVerts)
            for (int k = 0; k < 4; ++k) tetVerts[4 * t + k] = s[(code[(size_t)t] >> (2 * k)) & 3];
        if (tetCell && ctx->cellFromVertex) tetCell[t] = s[3] - ctx->nPoints;
    }
    return CPF_OK;
}

int cpf_mesh_download_neighbours(cpf_context *ctx, int *nbr)
{
    if (!ctx || !ctx->have_mesh || !nbr) return fail(ctx, CPF_ERR_INVALID, "no mesh uploaded");
    cudaSetDevice(ctx->device);
    std::vector<int4> l2((size_t)ctx->nTets * 2), l((size_t)ctx->nTets);
    std::vector<uint16_t> code((size_t)ctx->nTets);
    CPF_CUDA(ctx, cudaMemcpy(l2.data(), ctx->d_tetrec, sizeof(int4) * l2.size(), cudaMemcpyDeviceToHost));
    for (size_t t = 0; t < l.size(); ++t) l[t] = l2[2 * t];
    CPF_CUDA(ctx, cudaMemcpy(code.data(), ctx->d_tetcode, sizeof(uint16_t) * code.size(), cudaMemcpyDeviceToHost));
    for (long long t = 0; t < ctx->nTets; ++t) {
        const int lk[4] = { l[(size_t)t].x, l[(size_t)t].y, l[(size_t)t].z, l[(size_t)t].w };
        for (int k = 0; k < 4; ++k) {
            const int v = lk[(code[(size_t)t] >> (2 * k)) & 3];
            nbr[4 * t + k] = v < 0 ? v : (v >> 2);
        }
    }
    return CPF_OK;
}

// ---------------------------------------------------------------------------------------------
// flow field
// ---------------------------------------------------------------------------------------------
int cpf_update_velocity(cpf_context *ctx, const double *U, int on_device)
{
    if (!ctx || !ctx->have_mesh || !U) return fail(ctx, CPF_ERR_INVALID, "cpf_update_velocity: no mesh or null field");
    cudaSetDevice(ctx->device);
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->nCells;
    const unsigned grid = (unsigned)((ctx->nCells + 255) / 256);
    // double-buffered: sub-steps already enqueued keep reading the previous field
    const int nb = 1 - ctx->ucur;
    // every refresh marks the end of the kernels that read the buffer in use (they were all enqueued before this call)
    CPF_CUDA(ctx, cudaEventRecord(ctx->evRead[ctx->ucur], ctx->stream));
    if (on_device) {
        // device field (e.g. the NCCL broadcast buffer): repacked straight into the idle buffer on the compute stream;
        // the caller's buffer is free again when this kernel has run (stream order)
        k_pack_velocity<<<grid, 256, 0, ctx->stream>>>(ctx->nCells, U, ctx->d_ucell[nb]);
    } else {
        // Host field: upload and repack run on the copy stream, concurrently with the sub-steps already enqueued on the
        // compute stream (they read the other buffer).  The buffer being overwritten was last read by the kernels
        // enqueued before the PREVIOUS refresh (evRead[nb]); kernels enqueued from now on wait for the repack.
        if (!ctx->d_ustage) CPF_CUDA(ctx, cudaMalloc(&ctx->d_ustage, bytes));
        CPF_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->evRead[nb], 0));
        CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_ustage, U, bytes, cudaMemcpyHostToDevice, ctx->copyStream));
        k_pack_velocity<<<grid, 256, 0, ctx->copyStream>>>(ctx->nCells, ctx->d_ustage, ctx->d_ucell[nb]);
        CPF_CUDA(ctx, cudaEventRecord(ctx->evCopy, ctx->copyStream));
        CPF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evCopy, 0));
    }
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    ctx->ucur = nb;
    if (ctx->cfg.interp == CPF_INTERP_VERTEX && ctx->d_pc_off) return launch_point_interp(ctx);
    return CPF_OK;
}

} // extern "C"

namespace cpf {
// second half of a field refresh whose data already sits in a device staging buffer filled ON THE COPY STREAM
// (cpf_comm.cu: NCCL broadcast / slice exchange): repack there, then make the compute stream wait
int update_velocity_staged(cpf_context *ctx, const double *d_stage)
{
    const int nb = 1 - ctx->ucur;
    CPF_CUDA(ctx, cudaEventRecord(ctx->evRead[ctx->ucur], ctx->stream));
    CPF_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->evRead[nb], 0));
    k_pack_velocity<<<(unsigned)((ctx->nCells + 255) / 256), 256, 0, ctx->copyStream>>>(ctx->nCells, d_stage, ctx->d_ucell[nb]);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    CPF_CUDA(ctx, cudaEventRecord(ctx->evCopy, ctx->copyStream));
    CPF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evCopy, 0));
    ctx->ucur = nb;
    if (ctx->cfg.interp == CPF_INTERP_VERTEX && ctx->d_pc_off) return launch_point_interp(ctx);
    return CPF_OK;
}
} // namespace cpf

extern "C" {

int cpf_update_vertex_velocity(cpf_context *ctx, const double *Uvert, int on_device)
{
    if (!ctx || !ctx->have_mesh || !Uvert) return fail(ctx, CPF_ERR_INVALID, "cpf_update_vertex_velocity: no mesh or null field");
    cudaSetDevice(ctx->device);
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->nVerts;
    if (!ctx->d_uvert) CPF_CUDA(ctx, cudaMalloc(&ctx->d_uvert, bytes));
    CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_uvert, Uvert, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    return CPF_OK;
}

// ---------------------------------------------------------------------------------------------
// particles
// ---------------------------------------------------------------------------------------------
int cpf_set_particles(cpf_context *ctx, long long n, const double *xyzw)
{
    if (!ctx || n < 0 || (n > 0 && !xyzw)) return fail(ctx, CPF_ERR_INVALID, "cpf_set_particles: bad arguments");
    cudaSetDevice(ctx->device);
    int rc = alloc_particles(ctx, n);
    if (rc) return rc;
    if (n) CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_pos[0], xyzw, sizeof(double4) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CPF_OK;
}

int cpf_seed_box_slice(cpf_context *ctx, long long first, long long count, const double lo[3], const double hi[3], unsigned long long seed)
{
    if (!ctx || first < 0 || count < 0 || !lo || !hi) return fail(ctx, CPF_ERR_INVALID, "cpf_seed_box_slice: bad arguments");
    std::vector<double> p;
    try { p.resize((size_t)count * 4); } catch (const std::bad_alloc &) { return fail(ctx, CPF_ERR_NOMEM, "cpf_seed_box_slice: out of host memory"); }
    for (int ax = 0; ax < 3; ++ax) {
        const unsigned long long key = splitmix64(seed ^ ((unsigned long long)ax * 0xD1342543DE82EF95ull));
        for (long long k = 0; k < count; ++k) {
            const unsigned long long bits = splitmix64((unsigned long long)(first + k) * 0x2545F4914F6CDD1Dull + key); // keyed by GLOBAL index
            const double u = (double)(bits >> 11) * (1.0 / 9007199254740992.0);
            p[(size_t)k * 4 + ax] = lo[ax] + u * (hi[ax] - lo[ax]); // lower + r*(upper-lower), particles.cu:88-91
        }
    }
    for (long long k = 0; k < count; ++k) p[(size_t)k * 4 + 3] = 1.0;
    int rc = cpf_set_particles(ctx, count, p.data());
    if (rc) return rc;
    ctx->id_base = (unsigned long long)first;
    return CPF_OK;
}

int cpf_seed_box(cpf_context *ctx, long long n, const double lo[3], const double hi[3], unsigned long long seed)
{
    return cpf_seed_box_slice(ctx, 0, n, lo, hi, seed);
}

int cpf_set_tets(cpf_context *ctx, const int *tet)
{
    if (!ctx || !tet || ctx->n == 0) return fail(ctx, CPF_ERR_INVALID, "cpf_set_tets: no particles");
    if (ctx->permuted) return fail(ctx, CPF_ERR_INVALID, "cpf_set_tets after a sort is not supported");
    if (ctx->have_mesh)
        for (long long i = 0; i < ctx->n; ++i)
            if (tet[i] >= ctx->nTets) return fail(ctx, CPF_ERR_INVALID, "cpf_set_tets: tet[%lld] = %d, the mesh has %lld tets", i, tet[i], ctx->nTets);
    cudaSetDevice(ctx->device);
    CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_tet[ctx->pcur], tet, sizeof(int) * (size_t)ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_tets = true;
    return CPF_OK;
}

int cpf_locate_initial(cpf_context *ctx)
{
    if (!ctx) return CPF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return locate_particles(ctx);
}

int cpf_relocate_lost(cpf_context *ctx)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (!ctx->have_mesh || !ctx->have_tets) return fail(ctx, CPF_ERR_INVALID, "cpf_relocate_lost: mesh and located particles are required");
    cudaSetDevice(ctx->device);
    return locate_particles(ctx, true);
}

int cpf_reseed_inactive(cpf_context *ctx, const double lo[3], const double hi[3], unsigned long long seed, long long *nReseeded)
{
    if (!ctx || !lo || !hi) return fail(ctx, CPF_ERR_INVALID, "cpf_reseed_inactive: bad arguments");
    if (!ctx->have_mesh || !ctx->have_tets) return fail(ctx, CPF_ERR_INVALID, "cpf_reseed_inactive: mesh and located particles are required");
    if (nReseeded) *nReseeded = 0;
    if (ctx->n == 0) return CPF_OK;
    cudaSetDevice(ctx->device);
    int rc = CPF_OK;
    // the count of re-seeded particles: two spare words of the queue counters (the scratch buffer holds the location pass's workspace)
    unsigned long long *d_n = reinterpret_cast<unsigned long long *>(ctx->d_queue_count + 60);
    CPF_CUDA(ctx, cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), ctx->stream));
    const int a = ctx->pcur;
    k_reseed_inactive<<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n, ctx->d_pos[a], ctx->d_tet[a], ctx->d_pid[a], ctx->d_vel[a],
                                                                               make_double3(lo[0], lo[1], lo[2]), make_double3(hi[0], hi[1], hi[2]),
                                                                               splitmix64(seed ^ ctx->step_index), ctx->id_base, d_n);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    rc = locate_particles(ctx, true); // BVH location of exactly the particles just re-seeded (active, negative tet id)
    if (rc) return rc;
    ctx->statBaseValid = ctx->statScanQueued = false; // the counters no longer explain n_active: the next statistics request scans
    if (nReseeded) {
        unsigned long long h = 0;
        CPF_CUDA(ctx, cudaMemcpyAsync(&h, d_n, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *nReseeded = (long long)h;
    }
    return CPF_OK;
}

int cpf_init_rng(cpf_context *ctx)
{
    if (!ctx || ctx->n == 0) return fail(ctx, CPF_ERR_INVALID, "cpf_init_rng: no particles");
    cudaSetDevice(ctx->device);
    return launch_init_rng(ctx);
}

// ---------------------------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------------------------
int cpf_substeps(cpf_context *ctx, int n, double dt)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (ctx->n == 0 || n <= 0) return CPF_OK;
    if (!ctx->have_mesh || !ctx->have_tets) return fail(ctx, CPF_ERR_INVALID, "cpf_substeps: mesh and located particles are required");
    if ((ctx->cfg.integrator != CPF_EULER || ctx->cfg.interp != CPF_INTERP_TET) && ctx->cfg.locator != CPF_LOCATOR_CONVEX)
        return fail(ctx, CPF_ERR_INVALID, "RK2/RK4 and vertex interpolation are available with the convex locator only");
    if (ctx->cfg.integrator != CPF_EULER && ctx->cfg.integrator != CPF_RK2 && ctx->cfg.integrator != CPF_RK4)
        return fail(ctx, CPF_ERR_INVALID, "unknown integrator %d", ctx->cfg.integrator);
    if (ctx->cfg.interp == CPF_INTERP_VERTEX && !ctx->d_uvert)
        return fail(ctx, CPF_ERR_INVALID, "vertex interpolation: no vertex field (cpf_update_velocity after an upload with interp = VERTEX, or cpf_update_vertex_velocity)");
    cudaSetDevice(ctx->device);
    CPF_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    // cfg.fuse_substeps == 0: library default; the glue chunks at save boundaries, so fusing never skips an output
    const int fuse = std::min(ctx->cfg.fuse_substeps > 0 ? ctx->cfg.fuse_substeps : default_fused_substeps(ctx), max_fused_substeps(ctx));
    int done = 0;
    while (done < n) {
        if (ctx->cfg.sort_interval > 0 && ctx->since_sort >= ctx->cfg.sort_interval) {
            int rc = sort_particles_by_cell(ctx);
            if (rc) return rc;
        }
        // balanced chunks: 10 sub-steps with at most 8 per launch sequence run as 5 + 5, not 8 + 2
        const int left = n - done;
        int k = (left + (left + fuse - 1) / fuse - 1) / ((left + fuse - 1) / fuse);
        if (ctx->cfg.sort_interval > 0) k = std::min(k, std::max(1, ctx->cfg.sort_interval - ctx->since_sort));
        int rc = launch_substeps(ctx, k, dt, done + k == n);
        if (rc) return rc;
        done += k;
        ctx->since_sort += k;
    }
    CPF_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    return CPF_OK;
}

int cpf_advect(cpf_context *ctx, double deltaT, int *nCyclesOut)
{
    if (!ctx) return CPF_ERR_INVALID;
    // src/advect.H:36-37
    int nCycles = (int)std::max(std::ceil(deltaT / ctx->cfg.dt), 1.0);
    const double cycleDt = deltaT / nCycles;
    if (nCyclesOut) *nCyclesOut = nCycles;
    return cpf_substeps(ctx, nCycles, cycleDt);
}

int cpf_initial_advect(cpf_context *ctx)
{
    if (!ctx || !ctx->have_mesh || !ctx->have_tets) return fail(ctx, CPF_ERR_INVALID, "cpf_initial_advect: mesh and located particles are required");
    cudaSetDevice(ctx->device);
    return launch_initial_advect(ctx, ctx->cfg.dt);
}

int cpf_sort_particles(cpf_context *ctx)
{
    if (!ctx) return CPF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return sort_particles_by_cell(ctx);
}

int cpf_last_step_ms(cpf_context *ctx, float *ms)
{
    if (!ctx || !ms) return CPF_ERR_INVALID;
    CPF_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    CPF_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return CPF_OK;
}

// ---------------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------------
int cpf_download(cpf_context *ctx, double *xyzw, double *vel, int *tet)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (ctx->n == 0) return CPF_OK;
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->n;
    const int a = ctx->pcur;
    if (!ctx->permuted) {
        if (xyzw) CPF_CUDA(ctx, cudaMemcpyAsync(xyzw, ctx->d_pos[a], sizeof(double4) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (vel) CPF_CUDA(ctx, cudaMemcpyAsync(vel, ctx->d_vel[a], sizeof(double4) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (tet) CPF_CUDA(ctx, cudaMemcpyAsync(tet, ctx->d_tet[a], sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        int rc = ensure_scratch(ctx, n * (2 * sizeof(double4) + sizeof(int)));
        if (rc) return rc;
        double4 *dp = (double4 *)ctx->d_scratch, *dv = dp + n;
        int *dt = (int *)(dv + n);
        rc = gather_original_order(ctx, xyzw ? dp : nullptr, vel ? dv : nullptr, tet ? dt : nullptr);
        if (rc) return rc;
        if (xyzw) CPF_CUDA(ctx, cudaMemcpyAsync(xyzw, dp, sizeof(double4) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (vel) CPF_CUDA(ctx, cudaMemcpyAsync(vel, dv, sizeof(double4) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (tet) CPF_CUDA(ctx, cudaMemcpyAsync(tet, dt, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CPF_OK;
}

int cpf_download_cells(cpf_context *ctx, int *cell)
{
    if (!ctx || !cell || !ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "cpf_download_cells: bad arguments");
    if (ctx->n == 0) return CPF_OK;
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->n;
    int rc = ensure_scratch(ctx, sizeof(int) * n);
    if (rc) return rc;
    // tet -> cell on the device, scattered to ORIGINAL particle order (4 bytes per particle cross PCIe, not the tet table)
    k_particle_cells<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(mesh_view(ctx), ctx->n, ctx->d_tet[ctx->pcur], ctx->d_pid[ctx->pcur], (int *)ctx->d_scratch);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    CPF_CUDA(ctx, cudaMemcpyAsync(cell, ctx->d_scratch, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CPF_OK;
}

int cpf_stats_request(cpf_context *ctx, int full)
{
    if (!ctx) return CPF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return stats_request(ctx, full != 0);
}

int cpf_stats_collect(cpf_context *ctx, cpf_stats *out)
{
    if (!ctx || !out) return CPF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return stats_collect(ctx, out);
}

int cpf_stats_get(cpf_context *ctx, cpf_stats *out)
{
    if (!ctx || !out) return CPF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cpf_stats tmp;
    while (ctx->statHead != ctx->statTail) { // results nobody collected: drop them, oldest first
        int rc = stats_collect(ctx, &tmp);
        if (rc) return rc;
    }
    int rc = stats_request(ctx, true);
    if (rc) return rc;
    return stats_collect(ctx, out);
}

long long cpf_num_particles(cpf_context *ctx) { return ctx ? ctx->n : 0; }

int cpf_device_pointers(cpf_context *ctx, void **pos4, void **tet, void **ucell)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (pos4) *pos4 = ctx->d_pos[ctx->pcur];
    if (tet) *tet = ctx->d_tet[ctx->pcur];
    if (ucell) *ucell = ctx->d_ucell[ctx->ucur];
    return CPF_OK;
}

int cpf_debug_normals(cpf_context *ctx, int k, double *xi)
{
    if (!ctx || !xi || ctx->n == 0 || k < 1 || k > 4096) return fail(ctx, CPF_ERR_INVALID, "cpf_debug_normals: bad arguments");
    cudaSetDevice(ctx->device);
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->n * (size_t)k;
    int rc = ensure_scratch(ctx, bytes); // the context's scratch buffer: no allocation per call
    if (rc) return rc;
    rc = launch_debug_normals(ctx, k, (double *)ctx->d_scratch);
    if (rc) return rc;
    CPF_CUDA(ctx, cudaMemcpyAsync(xi, ctx->d_scratch, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CPF_OK;
}

int cpf_debug_next_normals(cpf_context *ctx, double *xi) { return cpf_debug_normals(ctx, 1, xi); }

long long cpf_launch_count(cpf_context *ctx) { return ctx ? ctx->launches : 0; }

// writeParticles2VTU (cuda/utils.cpp:144-283): same file name, arrays and ASCII layout so existing
// ParaView states keep working.  ParticleTetID carries the barycentric-mode id array of the
// reference (unused in the default build, Appendix A.9) -- both arrays hold the current tet here.
int cpf_write_vtu(cpf_context *ctx, const char *dir, unsigned step)
{
    if (!ctx) return CPF_ERR_INVALID;
    const long long n = ctx->n;
    std::vector<double> p((size_t)n * 4), v((size_t)n * 4);
    std::vector<int> tet((size_t)n);
    int rc = cpf_download(ctx, p.data(), v.data(), tet.data());
    if (rc) return rc;
    char name[1200];
    snprintf(name, sizeof name, "%s%sparticle_%04u.vtu", dir ? dir : "", (dir && *dir) ? "/" : "", step);
    FILE *fp = fopen(name, "w");
    if (!fp) return fail(ctx, CPF_ERR_INVALID, "cannot open %s", name);
    fprintf(fp, "<VTKFile type='UnstructuredGrid' version='1.0' byte_order='LittleEndian' header_type='UInt64'>\n<UnstructuredGrid>\n");
    fprintf(fp, "<Piece NumberOfCells='%lld' NumberOfPoints='%lld'>\n<Points>\n", n, n);
    fprintf(fp, "<DataArray NumberOfComponents='3' type='Float64' Name='Position' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "%.15lf %.15lf %.15lf\n", p[4 * i], p[4 * i + 1], p[4 * i + 2]);
    fprintf(fp, "</DataArray>\n</Points>\n<PointData>\n");
    fprintf(fp, "<DataArray NumberOfComponents='1' type='Int32' Name='ParticleType' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "%d\n", (int)p[4 * i + 3]);
    fprintf(fp, "</DataArray>\n<DataArray NumberOfComponents='1' type='Int32' Name='ParticleID' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "%lld\n", (long long)ctx->id_base + i); // global id (one rank per GPU: index ranges of one cloud)
    fprintf(fp, "</DataArray>\n<DataArray NumberOfComponents='1' type='Int32' Name='ParticleTetID' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "%d\n", tet[(size_t)i]);
    fprintf(fp, "</DataArray>\n<DataArray NumberOfComponents='1' type='Int32' Name='ConvexTetID' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "%d\n", tet[(size_t)i]);
    fprintf(fp, "</DataArray>\n<DataArray NumberOfComponents='3' type='Float32' Name='vels' format='ascii'>\n");
    double totalKE = 0.0;
    for (long long i = 0; i < n; ++i) {
        if (std::isnan(v[4 * i])) fprintf(fp, "%lf %lf %lf\n", 0.0, 0.0, 0.0);
        else fprintf(fp, "%lf %lf %lf\n", v[4 * i], v[4 * i + 1], v[4 * i + 2]);
    }
    fprintf(fp, "</DataArray>\n<DataArray NumberOfComponents='1' type='Float32' Name='KEs' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) {
        const double ke = 0.5 * (v[4 * i] * v[4 * i] + v[4 * i + 1] * v[4 * i + 1] + v[4 * i + 2] * v[4 * i + 2]);
        fprintf(fp, "%lf\n", ke);
        totalKE += ke;
    }
    fprintf(fp, "</DataArray>\n</PointData>\n<Cells>\n<DataArray type='Int32' Name='connectivity' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "%lld\n", i);
    fprintf(fp, "</DataArray>\n<DataArray type='Int32' Name='offsets' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "%lld\n", i + 1);
    fprintf(fp, "</DataArray>\n<DataArray type='UInt8' Name='types' format='ascii'>\n");
    for (long long i = 0; i < n; ++i) fprintf(fp, "1\n");
    fprintf(fp, "</DataArray>\n</Cells>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
    fclose(fp);
    printf("#adv: System Kinetic Energy=%lf\n", totalKE);
    return CPF_OK;
}

} // extern "C"
