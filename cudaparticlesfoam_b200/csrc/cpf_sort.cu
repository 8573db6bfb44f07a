// cpf_sort.cu -- sort-by-cell (locality), original-order gather for downloads, statistics.
//
// The reference never reorders particles (thread i == particle i forever,
// third_party/RTXAdvect/query/ConvexQuery.cu:146), so neighbouring threads gather unrelated tets.
// Here particles are periodically counting-sorted by the cell that contains them; `pid` carries the
// original index so that every download is returned in the reference's order.
#include <cub/cub.cuh>

#include <algorithm>

#include "cpf_internal.h"

namespace cpf {

CPF_DEV int sort_key(const MeshView &m, int tet)
{
    if (tet < 0) return m.nCells; // frozen / lost particles go last
    const int4 v = ld_int4(m.tetv, tet);
    return tet_cell(m, tet, v);
}

__global__ void k_hist(const MeshView m, const int *__restrict__ tet, long long n, int *__restrict__ hist)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + sort_key(m, tet[i]), 1);
}

__global__ void k_scatter(const MeshView m, long long n, int *__restrict__ cursor, const double4 *__restrict__ pos,
                          const int *__restrict__ tet, const int *__restrict__ pid, const double4 *__restrict__ vel,
                          const curandState_t *__restrict__ rng, double4 *__restrict__ pos2, int *__restrict__ tet2,
                          int *__restrict__ pid2, double4 *__restrict__ vel2, curandState_t *__restrict__ rng2)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = tet[i];
    // particles arrive nearly sorted: neighbouring lanes mostly share the key, so one atomic per (warp, key) instead of
    // one per particle; the lanes of a group take consecutive slots in lane order (keeps equal keys in their old order)
    const int key = sort_key(m, t);
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, key);
    const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cursor + key, __popc(grp));
    base = __shfl_sync(act, base, leader);
    const int dst = base + __popc(grp & ((1u << lane) - 1u));
    pos2[dst] = pos[i];
    tet2[dst] = t;
    pid2[dst] = pid[i];
    vel2[dst] = vel[i];
    if (rng) rng2[dst] = rng[i];
}

int sort_particles_by_cell(cpf_context *ctx)
{
    if (ctx->n == 0 || !ctx->have_mesh || !ctx->have_tets) return CPF_OK;
    cudaStream_t st = ctx->stream;
    const long long n = ctx->n;
    const int nb = (int)ctx->nCells + 1;
    if (!ctx->d_sort_hist) CPF_CUDA(ctx, cudaMalloc(&ctx->d_sort_hist, sizeof(int) * (size_t)(ctx->nCells + 2)));
    CPF_CUDA(ctx, cudaMemsetAsync(ctx->d_sort_hist, 0, sizeof(int) * (size_t)nb, st));
    const MeshView m = mesh_view(ctx);
    const int a = ctx->pcur, b = 1 - a;
    const unsigned grid = (unsigned)((n + 255) / 256);
    k_hist<<<grid, 256, 0, st>>>(m, ctx->d_tet[a], n, ctx->d_sort_hist);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_sort_hist, ctx->d_sort_hist, nb, st);
    int rc = ensure_scratch(ctx, tmp);
    if (rc) return rc;
    CPF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_scratch, tmp, ctx->d_sort_hist, ctx->d_sort_hist, nb, st));
    const bool useRng = ctx->rng_ready && ctx->d_rng[a] && ctx->d_rng[b];
    k_scatter<<<grid, 256, 0, st>>>(m, n, ctx->d_sort_hist, ctx->d_pos[a], ctx->d_tet[a], ctx->d_pid[a], ctx->d_vel[a],
                                    useRng ? ctx->d_rng[a] : nullptr, ctx->d_pos[b], ctx->d_tet[b], ctx->d_pid[b], ctx->d_vel[b],
                                    useRng ? ctx->d_rng[b] : nullptr);
    ctx->launches += 4;
    CPF_CUDA(ctx, cudaGetLastError());
    ctx->pcur = b;
    ctx->permuted = true;
    ctx->since_sort = 0;
    return CPF_OK;
}

__global__ void k_gather(long long n, const int *__restrict__ pid, const double4 *__restrict__ pos, const double4 *__restrict__ vel,
                         const int *__restrict__ tet, double4 *__restrict__ pos_o, double4 *__restrict__ vel_o, int *__restrict__ tet_o)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long o = pid[i];
    if (pos_o) pos_o[o] = pos[i];
    if (vel_o) vel_o[o] = vel[i];
    if (tet_o) tet_o[o] = tet[i];
}

int gather_original_order(cpf_context *ctx, double4 *d_pos_out, double4 *d_vel_out, int *d_tet_out)
{
    const int a = ctx->pcur;
    k_gather<<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->n, ctx->d_pid[a], ctx->d_pos[a], ctx->d_vel[a], ctx->d_tet[a],
                                                                        d_pos_out, d_vel_out, d_tet_out);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    return CPF_OK;
}

// cudaReportParticles (cuda/particles.cu:763-775) + kinetic energy (cuda/utils.cpp:253-258)
__global__ void __launch_bounds__(256) k_stats(long long n, const double4 *__restrict__ pos, const int *__restrict__ tet,
                                               const double4 *__restrict__ vel, unsigned long long *__restrict__ out_counts,
                                               double *__restrict__ out_ke)
{
    // grid-stride over a fixed grid, block reduction, ONE atomic per block and quantity
    unsigned long long active = 0, neg = 0;
    double ke = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        active += pos[i].w != 0.0;
        neg += tet[i] < 0;
        const double4 v = vel[i];
        ke += 0.5 * (v.x * v.x + v.y * v.y + v.z * v.z);
    }
    __shared__ unsigned long long sa[8], sn[8];
    __shared__ double sk[8];
    for (int o = 16; o > 0; o >>= 1) {
        active += __shfl_down_sync(0xffffffffu, active, o);
        neg += __shfl_down_sync(0xffffffffu, neg, o);
        ke += __shfl_down_sync(0xffffffffu, ke, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sa[wid] = active; sn[wid] = neg; sk[wid] = ke; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { active += sa[k]; neg += sn[k]; ke += sk[k]; }
        atomicAdd(out_counts, active);
        atomicAdd(out_counts + 1, neg);
        atomicAdd(out_ke, ke);
    }
}

// slot layout (fp64 on the device so that ONE ncclAllReduce(sum) reduces a slot across ranks; every count is < 2^53)
enum { ST_N = 0, ST_ACTIVE, ST_NEG, ST_ESC, ST_REFL, ST_EXACT, ST_HOPS, ST_SUB, ST_KE, ST_FRZ, ST_LOST, ST_FULL };

__global__ void k_stats_pack(long long n, const unsigned long long *__restrict__ counters, const unsigned long long *__restrict__ scan,
                             const double *__restrict__ scanKe, int full, double *__restrict__ out)
{
    if (threadIdx.x || blockIdx.x) return;
    out[ST_N] = (double)n;
    out[ST_ACTIVE] = full ? (double)scan[0] : 0.0;
    out[ST_NEG] = full ? (double)scan[1] : 0.0;
    out[ST_KE] = full ? *scanKe : 0.0;
    out[ST_ESC] = (double)counters[CNT_ESCAPED];
    out[ST_REFL] = (double)counters[CNT_REFLECT];
    out[ST_EXACT] = (double)counters[CNT_EXACT];
    out[ST_HOPS] = (double)counters[CNT_HOPS];
    out[ST_SUB] = (double)counters[CNT_SUBSTEPS];
    out[ST_FRZ] = (double)counters[CNT_FROZEN];
    out[ST_LOST] = (double)counters[CNT_LOST];
    out[ST_FULL] = full ? 1.0 : 0.0;
}

static int stats_init(cpf_context *ctx)
{
    if (ctx->h_stat) return CPF_OK;
    CPF_CUDA(ctx, cudaMallocHost(&ctx->h_stat, sizeof(double) * cpf_context::STAT_SLOTS * cpf_context::STAT_WORDS));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_stat, sizeof(double) * cpf_context::STAT_WORDS * cpf_context::STAT_SLOTS));
    for (int k = 0; k < cpf_context::STAT_SLOTS; ++k) CPF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evStat[k], cudaEventDisableTiming));
    return CPF_OK;
}

void stats_release(cpf_context *ctx)
{
    if (ctx->h_stat) cudaFreeHost(ctx->h_stat);
    cudaFree(ctx->d_stat);
    for (int k = 0; k < cpf_context::STAT_SLOTS; ++k) if (ctx->evStat[k]) cudaEventDestroy(ctx->evStat[k]);
    ctx->h_stat = nullptr; ctx->d_stat = nullptr;
}

// Enqueue one statistics read-back behind everything submitted so far; no host synchronisation.  full: scan the particle
// arrays (exact n_active / n_negative_tet / kinetic energy at that point); light: the cumulative counters only (64 bytes).
// With a communicator (cpf_comm_init) the slot is summed over the ranks on the device before it crosses PCIe.
int stats_request(cpf_context *ctx, bool full)
{
    int rc = stats_init(ctx);
    if (rc) return rc;
    if (ctx->statHead - ctx->statTail >= cpf_context::STAT_SLOTS) return fail(ctx, CPF_ERR_INVALID, "cpf_stats_request: %d requests outstanding, collect one first", cpf_context::STAT_SLOTS);
    if (!ctx->statBaseValid && !ctx->statScanQueued) full = true; // nothing to derive the light numbers from yet
    if (full) ctx->statScanQueued = true;                        // results are collected in order: the scan arrives first
    const int slot = ctx->statHead % cpf_context::STAT_SLOTS;
    cudaStream_t st = ctx->stream;
    rc = ensure_scratch(ctx, 64);
    if (rc) return rc;
    unsigned long long *d_c = (unsigned long long *)ctx->d_scratch;
    double *d_ke = (double *)(d_c + 2);
    if (full) {
        CPF_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 64, st));
        if (ctx->n) {
            const int a = ctx->pcur;
            k_stats<<<(unsigned)std::min<long long>((ctx->n + 255) / 256, 148 * 8), 256, 0, st>>>(ctx->n, ctx->d_pos[a], ctx->d_tet[a], ctx->d_vel[a], d_c, d_ke);
            ctx->launches++;
        }
    }
    double *d_slot = ctx->d_stat + slot * cpf_context::STAT_WORDS;
    k_stats_pack<<<1, 32, 0, st>>>(ctx->n, ctx->d_counters, d_c, d_ke, full ? 1 : 0, d_slot);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    rc = comm_reduce_stats(ctx, d_slot, cpf_context::STAT_WORDS);
    if (rc) return rc;
    CPF_CUDA(ctx, cudaMemcpyAsync(ctx->h_stat + slot * cpf_context::STAT_WORDS, d_slot, sizeof(double) * cpf_context::STAT_WORDS, cudaMemcpyDeviceToHost, st));
    CPF_CUDA(ctx, cudaEventRecord(ctx->evStat[slot], st));
    ctx->statFull[slot] = full;
    ctx->statHead++;
    return CPF_OK;
}

// Wait for the OLDEST outstanding request only (a request made one step ago has long arrived: no pipeline bubble).
int stats_collect(cpf_context *ctx, cpf_stats *out)
{
    memset(out, 0, sizeof *out);
    if (ctx->statHead == ctx->statTail) return fail(ctx, CPF_ERR_INVALID, "cpf_stats_collect: no outstanding cpf_stats_request");
    const int slot = ctx->statTail % cpf_context::STAT_SLOTS;
    CPF_CUDA(ctx, cudaEventSynchronize(ctx->evStat[slot]));
    ctx->statTail++;
    const double *h = reinterpret_cast<const double *>(ctx->h_stat) + slot * cpf_context::STAT_WORDS;
    out->n_particles = (long long)h[ST_N];
    out->n_escaped = (long long)h[ST_ESC];
    out->n_reflections = (long long)h[ST_REFL];
    out->n_exact = (long long)h[ST_EXACT];
    out->n_hops = (long long)h[ST_HOPS];
    out->n_substeps = (long long)h[ST_SUB];
    const long long frz = (long long)h[ST_FRZ];
    if (h[ST_FULL] != 0.0) {
        out->n_active = (long long)h[ST_ACTIVE];
        out->n_negative_tet = (long long)h[ST_NEG];
        out->kinetic_energy = h[ST_KE];
        ctx->statBaseValid = true;
        ctx->statBaseActive = out->n_active; ctx->statBaseNeg = out->n_negative_tet; ctx->statBaseKe = out->kinetic_energy;
        ctx->statBaseEsc = out->n_escaped; ctx->statBaseFrz = frz;
    } else {
        // light: every particle that stopped being active since the last scan was counted as an escape or as a freeze
        out->n_active = ctx->statBaseActive - (out->n_escaped - ctx->statBaseEsc) - (frz - ctx->statBaseFrz);
        out->n_negative_tet = ctx->statBaseNeg;
        out->kinetic_energy = ctx->statBaseKe;
    }
    out->reserved[0] = h[ST_FULL];
    return CPF_OK;
}

} // namespace cpf
