// cpf_locate.cu -- initial / lost-particle point location with a hand-written BVH.
//
// Replaces the reference's init-time seeding: OptiX shared-face triangle BVH + fp32 ray cast
// (third_party/RTXAdvect/optix/OptixTetQuery.cpp:53-108,143-271, optix/optixQueryKernel.cu:63-124)
// followed by the fp64 narrow phase baryQuery (query/RTQuery.cu:189-218, 295-310).
//
// Structure: tets sorted along a 63-bit Morton curve of their centroids; an implicit 8-ary tree of
// float boxes (rounded outward) over that order, level 0 = groups of 8 tets.  The top levels
// (<= 1024-node level and everything above it) are staged in shared memory, once per CTA of a resident grid; the lower
// levels and the candidate tets come through L2.  A particle is assigned the LOWEST tet id whose
// four reference barycentric coordinates (cuda/DeviceTetMesh.cuh:108-156) are all >= 0 -- the same
// answer as a brute-force scan, independent of traversal order; from a containing tet the
// reference's narrow phase returns immediately, so no further walk is needed.
#include <cub/cub.cuh>

#include "cpf_internal.h"

namespace cpf {

CPF_DEV unsigned long long spread21(unsigned long long x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(long long nTets, const int4 *__restrict__ tetv, const double4 *__restrict__ vpos, double3 lo,
                         double3 inv, unsigned long long *__restrict__ keys, int *__restrict__ ids)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTets) return;
    const int4 v = tetv[t];
    const D3 a = ld_vertex(vpos, v.x), b = ld_vertex(vpos, v.y), c = ld_vertex(vpos, v.z), d = ld_vertex(vpos, v.w);
    const double cx = 0.25 * (a.x + b.x + c.x + d.x), cy = 0.25 * (a.y + b.y + c.y + d.y), cz = 0.25 * (a.z + b.z + c.z + d.z);
    const double s = 2097151.0;
    const unsigned long long ix = (unsigned long long)fmin(fmax((cx - lo.x) * inv.x * s, 0.0), s);
    const unsigned long long iy = (unsigned long long)fmin(fmax((cy - lo.y) * inv.y * s, 0.0), s);
    const unsigned long long iz = (unsigned long long)fmin(fmax((cz - lo.z) * inv.z * s, 0.0), s);
    keys[t] = spread21(ix) | (spread21(iy) << 1) | (spread21(iz) << 2);
    ids[t] = (int)t;
}

// level 0 (groups of 8 tets in Morton order) and, below it, the box of every single tet in the same order (tlo/thi)
__global__ void k_bvh_leaves(long long nGroups, long long nTets, const int *__restrict__ order, const int4 *__restrict__ tetv,
                             const double4 *__restrict__ vpos, float4 *__restrict__ blo, float4 *__restrict__ bhi,
                             float4 *__restrict__ tlo, float4 *__restrict__ thi)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nGroups) return;
    double lx = 1e300, ly = 1e300, lz = 1e300, hx = -1e300, hy = -1e300, hz = -1e300;
    for (int q = 0; q < 8; ++q) {
        const long long i = 8 * g + q;
        if (i >= nTets) break;
        const int4 v = tetv[order[i]];
        const int ids[4] = { v.x, v.y, v.z, v.w };
        double tlx = 1e300, tly = 1e300, tlz = 1e300, thx = -1e300, thy = -1e300, thz = -1e300;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const D3 p = ld_vertex(vpos, ids[k]);
            tlx = fmin(tlx, p.x); tly = fmin(tly, p.y); tlz = fmin(tlz, p.z);
            thx = fmax(thx, p.x); thy = fmax(thy, p.y); thz = fmax(thz, p.z);
        }
        tlo[i] = make_float4(__double2float_rd(tlx), __double2float_rd(tly), __double2float_rd(tlz), 0.f);
        thi[i] = make_float4(__double2float_ru(thx), __double2float_ru(thy), __double2float_ru(thz), 0.f);
        lx = fmin(lx, tlx); ly = fmin(ly, tly); lz = fmin(lz, tlz);
        hx = fmax(hx, thx); hy = fmax(hy, thy); hz = fmax(hz, thz);
    }
    blo[g] = make_float4(__double2float_rd(lx), __double2float_rd(ly), __double2float_rd(lz), 0.f);
    bhi[g] = make_float4(__double2float_ru(hx), __double2float_ru(hy), __double2float_ru(hz), 0.f);
}

__global__ void k_bvh_up(long long nParents, long long nChildren, const float4 *__restrict__ clo, const float4 *__restrict__ chi,
                         float4 *__restrict__ plo, float4 *__restrict__ phi)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nParents) return;
    float4 lo = make_float4(3e38f, 3e38f, 3e38f, 0.f), hi = make_float4(-3e38f, -3e38f, -3e38f, 0.f);
    for (int q = 0; q < 8; ++q) {
        const long long i = 8 * g + q;
        if (i >= nChildren) break;
        const float4 a = clo[i], b = chi[i];
        lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
        hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
    }
    plo[g] = lo; phi[g] = hi;
}

#define CPF_BVH_MAX_LEVELS 12
struct BvhView {
    const float4 *lo[CPF_BVH_MAX_LEVELS];
    const float4 *hi[CPF_BVH_MAX_LEVELS];
    long long n[CPF_BVH_MAX_LEVELS];
    int nLevels;
    int topFirst;                       // first level staged in shared memory
    int topOffset[CPF_BVH_MAX_LEVELS];  // offset of each staged level in the smem arrays
    int topNodes;
    const float4 *topLo, *topHi;
    const int *order;
    const float4 *tetLo, *tetHi; // box of every tet, in `order`
    long long nTets;
};

// P rounded down / up to float: a float box (rounded outward at build time) contains P iff it overlaps [Plo, Phi] --
// the test never rejects a box that really contains the point
struct PF { float lx, ly, lz, hx, hy, hz; };
CPF_DEV bool in_box(const float4 lo, const float4 hi, const PF &p)
{
    return p.hx >= lo.x && p.lx <= hi.x && p.hy >= lo.y && p.ly <= hi.y && p.hz >= lo.z && p.lz <= hi.z;
}

// Work list of a location pass: key = 30-bit Morton code of the position for the particles to locate (lostOnly: active but
// carrying a negative tet id -- just re-seeded, or lost after a failed reflection sequence, which the reference freezes
// forever, cuda/particles.cu:334-338; else: every active particle), 0xffffffff for the others; sorted by key, the list
// puts neighbouring points into neighbouring lanes, which then walk the same nodes of the tree (lines shared in L1
// instead of 8 KB of boxes and vertices per point through L2), and the particles to skip at its end.
CPF_DEV unsigned spread10(unsigned x)
{
    x &= 0x3ffu;
    x = (x | x << 16) & 0x030000ffu;
    x = (x | x << 8) & 0x0300f00fu;
    x = (x | x << 4) & 0x030c30c3u;
    x = (x | x << 2) & 0x09249249u;
    return x;
}
__global__ void k_query_keys(const ParticleView pv, const int lostOnly, double3 lo, double3 inv, unsigned *__restrict__ keys,
                             int *__restrict__ ids, unsigned *__restrict__ count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    if (i < pv.n) {
        const double4 p = pv.pos[i];
        want = p.w != 0.0 && (!lostOnly || pv.tet[i] < 0);
        if (!lostOnly && p.w == 0.0) pv.tet[i] = -1;
        unsigned key = 0xffffffffu;
        if (want) {
            const unsigned ix = (unsigned)fmin(fmax((p.x - lo.x) * inv.x * 1023.0, 0.0), 1023.0);
            const unsigned iy = (unsigned)fmin(fmax((p.y - lo.y) * inv.y * 1023.0, 0.0), 1023.0);
            const unsigned iz = (unsigned)fmin(fmax((p.z - lo.z) * inv.z * 1023.0, 0.0), 1023.0);
            key = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
        }
        keys[i] = key;
        ids[i] = (int)i;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask && (threadIdx.x & 31) == 0) atomicAdd(count, (unsigned)__popc(mask));
}

// exact containment test of tet t (the reference's barycentric coordinates, cuda/DeviceTetMesh.cuh:108-156)
CPF_DEV bool tet_contains(const MeshView &m, int t, const D3 &P)
{
    Tet T;
    const int4 v = ld_int4(m.tetv, t);
    T.P[0] = ld_vertex(m.vpos, v.x); T.P[1] = ld_vertex(m.vpos, v.y); T.P[2] = ld_vertex(m.vpos, v.z); T.P[3] = ld_vertex(m.vpos, v.w);
    T.code = m.tetcode[t];
    double w[4];
    bary_exact(T, P, w);
    return w[0] >= 0.0 && w[1] >= 0.0 && w[2] >= 0.0 && w[3] >= 0.0;
}

#define CPF_LOCATE_CAND 24
// One resident wave of CTAs, each staging the top of the tree once and then taking particles in a grid-stride loop
// (the first *count entries of list, see k_query_keys).
// Phase 1, boxes only: depth-first traversal in which EVERY step is the same code for every lane -- pop a node, test the
// eight boxes below it (contiguous; the boxes of the single tets are the level below the leaf groups), push the
// children / note the tets that contain the point.  Phase 2: the exact test of the noted tets, all lanes in step.
// Keeping the fp64 tests out of the traversal loop is what keeps the lanes of a warp together: with them inside, a warp
// ran the exact test once per loop iteration for whichever lane happened to sit at a leaf.
__global__ void __launch_bounds__(128) k_locate(const MeshView m, const __grid_constant__ BvhView bv, const ParticleView pv,
                                                const int *__restrict__ list, const unsigned *__restrict__ count, const int lostOnly,
                                                unsigned long long *__restrict__ relocated)
{
    extern __shared__ float4 smem[];
    __shared__ const float4 *sPtrLo[CPF_BVH_MAX_LEVELS + 1], *sPtrHi[CPF_BVH_MAX_LEVELS + 1]; // [0]: the tets, [l + 1]: level l
    __shared__ long long sCnt[CPF_BVH_MAX_LEVELS + 1];
    const long long total = (long long)*count;
    if ((long long)blockIdx.x * blockDim.x >= total) return; // nothing for this CTA: do not stage either
    float4 *sLo = smem, *sHi = smem + bv.topNodes;
    for (int q = threadIdx.x; q < bv.topNodes; q += blockDim.x) { sLo[q] = bv.topLo[q]; sHi[q] = bv.topHi[q]; }
    if (threadIdx.x == 0) {
        sPtrLo[0] = bv.tetLo; sPtrHi[0] = bv.tetHi; sCnt[0] = bv.nTets;
        for (int l = 0; l < bv.nLevels; ++l) {
            const bool top = l >= bv.topFirst; // staged levels are read through their (generic) shared-memory address
            sPtrLo[l + 1] = top ? sLo + bv.topOffset[l] : bv.lo[l];
            sPtrHi[l + 1] = top ? sHi + bv.topOffset[l] : bv.hi[l];
            sCnt[l + 1] = bv.n[l];
        }
    }
    __syncthreads();
    unsigned found = 0;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        const long long i = (long long)list[w];
        const double4 p4 = pv.pos[i];
        const D3 P{ p4.x, p4.y, p4.z };
        const PF pf{ __double2float_rd(P.x), __double2float_rd(P.y), __double2float_rd(P.z),
                     __double2float_ru(P.x), __double2float_ru(P.y), __double2float_ru(P.z) };
        int best = 0x7fffffff;
        int stack[64];
        int cand[CPF_LOCATE_CAND];
        int sp = 0, nc = 0;
        const int root = bv.nLevels; // in the shifted numbering
        if (in_box(sPtrLo[root][0], sPtrHi[root][0], pf)) stack[sp++] = (root << 26) | 0;
        while (sp > 0) {
            const int e = stack[--sp];
            const int C = (e >> 26) - 1; // the level of the children
            const long long c0 = 8ll * (e & 0x3ffffff), n = sCnt[C];
            const float4 *lo = sPtrLo[C] + c0, *hi = sPtrHi[C] + c0;
            float4 a[8], b[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { // all sixteen requests go out before the first test
                const bool have = c0 + q < n;
                a[q] = have ? lo[q] : make_float4(1.f, 1.f, 1.f, 0.f);
                b[q] = have ? hi[q] : make_float4(0.f, 0.f, 0.f, 0.f); // an empty box
            }
            unsigned hit = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) hit |= in_box(a[q], b[q], pf) ? 1u << q : 0u;
            if (C > 0) {
                for (; hit; hit &= hit - 1u) { // (any order: every containing box is visited)
                    const int q = __ffs((int)hit) - 1;
                    if (sp < 64) stack[sp++] = (C << 26) | (int)(c0 + q);
                }
            } else {
                for (; hit; hit &= hit - 1u) {
                    const int t = __ldg(bv.order + c0 + (__ffs((int)hit) - 1));
                    if (nc < CPF_LOCATE_CAND) cand[nc++] = t;
                    else if (t < best && tet_contains(m, t, P)) best = t; // more candidates than the list holds: test at once
                }
            }
        }
        for (int k = 0; k < nc; ++k) {
            const int t = cand[k];
            if (t < best && tet_contains(m, t, P)) best = t;
        }
        if (lostOnly) {
            if (best != 0x7fffffff) { pv.tet[i] = best; ++found; }
        } else pv.tet[i] = (best == 0x7fffffff) ? -1 : best;
    }
    if (lostOnly) {
        found = __reduce_add_sync(0xffffffffu, found);
        if ((threadIdx.x & 31) == 0 && found) atomicAdd(relocated, (unsigned long long)found);
    }
}

void free_bvh(cpf_context *ctx)
{
    cudaFree(ctx->d_bvh_tet); ctx->d_bvh_tet = nullptr;
    cudaFree(ctx->d_bvh_tet_lo); cudaFree(ctx->d_bvh_tet_hi); ctx->d_bvh_tet_lo = ctx->d_bvh_tet_hi = nullptr;
    for (auto &l : ctx->bvh) { cudaFree(l.lo); cudaFree(l.hi); }
    ctx->bvh.clear();
    cudaFree(ctx->d_bvh_top_lo); cudaFree(ctx->d_bvh_top_hi);
    ctx->d_bvh_top_lo = ctx->d_bvh_top_hi = nullptr;
    ctx->bvh_top_offsets.clear();
}

int build_bvh(cpf_context *ctx)
{
    free_bvh(ctx);
    cudaStream_t st = ctx->stream;
    const long long nT = ctx->nTets;
    struct Tmp { // freed on every return path
        unsigned long long *k = nullptr, *k2 = nullptr;
        int *id = nullptr;
        void *tmp = nullptr;
        ~Tmp() { cudaFree(k); cudaFree(k2); cudaFree(id); cudaFree(tmp); }
    } T;
    CPF_CUDA(ctx, cudaMalloc(&T.k, sizeof(unsigned long long) * (size_t)nT));
    CPF_CUDA(ctx, cudaMalloc(&T.k2, sizeof(unsigned long long) * (size_t)nT));
    CPF_CUDA(ctx, cudaMalloc(&T.id, sizeof(int) * (size_t)nT));
    unsigned long long *d_k = T.k, *d_k2 = T.k2;
    int *d_id = T.id;
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_bvh_tet, sizeof(int) * (size_t)nT));
    double3 lo = make_double3(ctx->bbox_lo[0], ctx->bbox_lo[1], ctx->bbox_lo[2]);
    double3 inv;
    inv.x = 1.0 / fmax(ctx->bbox_hi[0] - ctx->bbox_lo[0], 1e-300);
    inv.y = 1.0 / fmax(ctx->bbox_hi[1] - ctx->bbox_lo[1], 1e-300);
    inv.z = 1.0 / fmax(ctx->bbox_hi[2] - ctx->bbox_lo[2], 1e-300);
    k_morton<<<(unsigned)((nT + 255) / 256), 256, 0, st>>>(nT, ctx->d_tetv, ctx->d_vpos, lo, inv, d_k, d_id);
    size_t tmpBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, d_k, d_k2, d_id, ctx->d_bvh_tet, nT, 0, 63, st);
    CPF_CUDA(ctx, cudaMalloc(&T.tmp, tmpBytes));
    CPF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(T.tmp, tmpBytes, d_k, d_k2, d_id, ctx->d_bvh_tet, nT, 0, 63, st));
    ctx->launches += 5;

    long long n = (nT + 7) / 8;
    BvhLevel L0{ nullptr, nullptr, n };
    CPF_CUDA(ctx, cudaMalloc(&L0.lo, sizeof(float4) * (size_t)n));
    CPF_CUDA(ctx, cudaMalloc(&L0.hi, sizeof(float4) * (size_t)n));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_bvh_tet_lo, sizeof(float4) * (size_t)nT));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_bvh_tet_hi, sizeof(float4) * (size_t)nT));
    k_bvh_leaves<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, nT, ctx->d_bvh_tet, ctx->d_tetv, ctx->d_vpos, L0.lo, L0.hi, ctx->d_bvh_tet_lo, ctx->d_bvh_tet_hi);
    ctx->launches++;
    ctx->bvh.push_back(L0);
    while (n > 1) {
        const long long np = (n + 7) / 8;
        BvhLevel L{ nullptr, nullptr, np };
        CPF_CUDA(ctx, cudaMalloc(&L.lo, sizeof(float4) * (size_t)np));
        CPF_CUDA(ctx, cudaMalloc(&L.hi, sizeof(float4) * (size_t)np));
        const BvhLevel &c = ctx->bvh.back();
        k_bvh_up<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(np, n, c.lo, c.hi, L.lo, L.hi);
        ctx->launches++;
        ctx->bvh.push_back(L);
        n = np;
    }
    if ((int)ctx->bvh.size() > CPF_BVH_MAX_LEVELS) return fail(ctx, CPF_ERR_INVALID, "BVH too deep");
    // top levels (first level with <= 1024 nodes and everything above) -> one contiguous array
    int first = 0;
    while (first < (int)ctx->bvh.size() - 1 && ctx->bvh[first].n > 1024) ++first;
    ctx->bvh_top_first_level = first;
    ctx->bvh_top_offsets.assign(ctx->bvh.size(), 0);
    int total = 0;
    for (int l = first; l < (int)ctx->bvh.size(); ++l) { ctx->bvh_top_offsets[l] = total; total += (int)ctx->bvh[l].n; }
    ctx->bvh_top_nodes = total;
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_bvh_top_lo, sizeof(float4) * (size_t)total));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_bvh_top_hi, sizeof(float4) * (size_t)total));
    for (int l = first; l < (int)ctx->bvh.size(); ++l) {
        CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_bvh_top_lo + ctx->bvh_top_offsets[l], ctx->bvh[l].lo, sizeof(float4) * (size_t)ctx->bvh[l].n,
                                      cudaMemcpyDeviceToDevice, st));
        CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_bvh_top_hi + ctx->bvh_top_offsets[l], ctx->bvh[l].hi, sizeof(float4) * (size_t)ctx->bvh[l].n,
                                      cudaMemcpyDeviceToDevice, st));
    }
    CPF_CUDA(ctx, cudaStreamSynchronize(st));
    return CPF_OK;
}

int locate_particles(cpf_context *ctx, bool lostOnly)
{
    if (!ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "no mesh uploaded");
    if (ctx->n == 0) return CPF_OK;
    BvhView bv{};
    bv.nLevels = (int)ctx->bvh.size();
    for (int l = 0; l < bv.nLevels; ++l) { bv.lo[l] = ctx->bvh[l].lo; bv.hi[l] = ctx->bvh[l].hi; bv.n[l] = ctx->bvh[l].n; bv.topOffset[l] = ctx->bvh_top_offsets[l]; }
    bv.topFirst = ctx->bvh_top_first_level;
    bv.topNodes = ctx->bvh_top_nodes;
    bv.topLo = ctx->d_bvh_top_lo; bv.topHi = ctx->d_bvh_top_hi;
    bv.order = ctx->d_bvh_tet;
    bv.tetLo = ctx->d_bvh_tet_lo; bv.tetHi = ctx->d_bvh_tet_hi;
    bv.nTets = ctx->nTets;
    const size_t smem = sizeof(float4) * 2 * (size_t)bv.topNodes;
    CPF_CUDA(ctx, cudaFuncSetAttribute(k_locate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSm = 0, sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_locate, 128, smem) != cudaSuccess || perSm < 1) perSm = 1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device) != cudaSuccess || sms < 1) sms = 148;
    const unsigned grid = (unsigned)std::min<long long>((ctx->n + 127) / 128, (long long)sms * perSm);
    const ParticleView pv = particle_view(ctx);
    // the work list: keys | sorted keys in the first deferral queue, ids | sorted ids in the second (no sub-steps are in flight
    // on this stream; a queue holds 2 n words), its length in the last queue counter word, CUB's workspace in the scratch buffer
    const long long n = ctx->n;
    unsigned *keys = reinterpret_cast<unsigned *>(ctx->d_queue[0]), *keys2 = keys + n;
    int *ids = reinterpret_cast<int *>(ctx->d_queue[1]), *ids2 = ids + n;
    unsigned *count = ctx->d_queue_count + 63;
    size_t tmpBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keys2, ids, ids2, n, 0, 32, ctx->stream);
    { int rc = ensure_scratch(ctx, tmpBytes); if (rc) return rc; }
    CPF_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(unsigned), ctx->stream));
    double3 lo = make_double3(ctx->bbox_lo[0], ctx->bbox_lo[1], ctx->bbox_lo[2]), inv;
    inv.x = 1.0 / fmax(ctx->bbox_hi[0] - ctx->bbox_lo[0], 1e-300);
    inv.y = 1.0 / fmax(ctx->bbox_hi[1] - ctx->bbox_lo[1], 1e-300);
    inv.z = 1.0 / fmax(ctx->bbox_hi[2] - ctx->bbox_lo[2], 1e-300);
    k_query_keys<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(pv, lostOnly ? 1 : 0, lo, inv, keys, ids, count);
    CPF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->d_scratch, tmpBytes, keys, keys2, ids, ids2, n, 0, 32, ctx->stream));
    k_locate<<<grid, 128, smem, ctx->stream>>>(mesh_view(ctx), bv, pv, ids2, count, lostOnly ? 1 : 0, ctx->d_counters + CNT_LOST);
    ctx->launches += 6;
    CPF_CUDA(ctx, cudaGetLastError());
    if (!lostOnly) ctx->have_tets = true;
    return CPF_OK;
}

} // namespace cpf
